// SIMT kernels of the SuperPoint + SuperGlue matcher (launch wrappers). fp32, channel-last.
#pragma once
#include "common.cuh"

namespace mfr {

int sp_conv1a(const float* img, const float* w /*[64][9]*/, const float* bias, float* out, int N, int H, int W,
              cudaStream_t st);
int maxpool2x2(const float* in, float* out, int N, int H, int W, int C, cudaStream_t st);
// softmax over 65 logits (row stride ldc), drop dustbin, depth-to-space -> scores [N][8h][8w]
int sp_score_map(const float* logits, int ldc, float* scores, int N, int h, int w, cudaStream_t st);
// simple_nms (superpoint.py:47-62): mask (0/1 floats); t0..t3 scratch of the same size
int sp_nms(const float* scores, float* mask, float* t0, float* t1, float* t2, float* t3, int N, int H, int W, int r,
           cudaStream_t st);
// threshold + border + top-k (superpoint.py:170-187) -> kpts [N][max_kp][2] (x,y), kscores, counts
int sp_select_keypoints(const float* scores, const float* mask, int N, int H, int W, float thr, int border,
                        int max_kp, float* kpts, float* kscores, int* counts, cudaStream_t st);
int l2norm256(float* x, size_t rows, cudaStream_t st);
int sp_sample_descriptors(const float* dense, const float* kpts, const int* counts, int N, int h, int w, int max_kp,
                          float* out, long long ldo, cudaStream_t st);
// SuperGlue
int sg_kenc_input(const float* kpts, const float* kscores, float* out, int total, int H, int W, cudaStream_t st);
// images are stored pair-interleaved (2p, 2p+1); cross != 0: image i attends to image i^1
int sg_attention(const float* qkv, long long ld, const int* counts, int cross, int max_kp, int n_img, float* out,
                 long long ldo, cudaStream_t st);
// tensor-core attention helpers (see spsg_ops.cu): in-place row softmax of the [img][4][K][K] scores, V transpose
int sg_softmax_rows(float* S, const int* counts, int cross, int max_kp, int n_img, cudaStream_t st);
int sg_transpose_v(const float* qkv, long long ld, float* vt, int max_kp, int n_img, cudaStream_t st);
int sg_sinkhorn(const float* S, int ldS, const int* counts, float alpha, int iters, int max_kp, int n_pairs,
                float* u, float* v, cudaStream_t st);
int sg_extract_matches(const float* S, int ldS, const int* counts, const float* u, const float* v, float thr,
                       const float* kpts, int max_kp, float* rmax, int* rarg, int* carg, int* matches0,
                       float* mscores0, float* mk0, float* mk1, int* m_offsets, int n_pairs, cudaStream_t st);

}  // namespace mfr
