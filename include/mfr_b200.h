/* mfr_b200 — C ABI of the B200-native feature-matching + scale-from-depth hot path.
 *
 * Drop-in boundary for nianticlabs/map-free-reloc (reference paths are relative to that repo):
 * the reference has no FFI of its own (pure Python over torch / OpenCV wheels), so each entry
 * point below names the reference call it replaces. All pointers are DEVICE pointers unless
 * a parameter is documented as host memory; `stream` is a cudaStream_t passed as void*.
 * Every function returns MFR_OK (0) or an MFR_ERR_* code; geometric failures are reported the
 * way the reference does it: NaN-filled R/t and an inlier count of 0
 * (lib/models/matching/pose_solver.py:30-33).
 */
#ifndef MFR_B200_H_
#define MFR_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define MFR_OK 0
#define MFR_ERR_CUDA 1
#define MFR_ERR_ARG 2
#define MFR_ERR_UNSUPPORTED 3

/* activation codes for the fused conv/GEMM epilogue */
#define MFR_ACT_NONE 0
#define MFR_ACT_RELU 1
#define MFR_ACT_LEAKY 2 /* slope 0.01: nn.LeakyReLU() default, resnet_fpn.py:74,81 */
#define MFR_ACT_ELU1 3  /* elu(x)+1: linear_attention.py:10-11 */

/* Library / device probe. Returns the number of SMs of the current device (0 if no device). */
int mfr_version(void);
int mfr_device_sm_count(void);

/* Implicit-GEMM convolution on tcgen05 tensor cores (TF32 inputs, fp32 accumulate), NHWC.
 * Replaces torch's cuDNN conv2d / cuBLAS linear / einsum calls of the matcher networks:
 *   etc/feature_matching_baselines/LoFTR/src/loftr/backbone/resnet_fpn.py:5-12 (conv1x1/conv3x3),
 *   .../loftr_module/transformer.py:22-32 (Linear), .../utils/coarse_matching.py:113 (einsum).
 *   x   [N][H][W][Cin]  pixel stride ldx floats        w [Cout][R*S][Cin]
 *   out [N][Ho][Wo][Cout] pixel stride ldo floats; residual like out with stride ldr (or NULL)
 *   b_batched != 0: w is [N][Cout][Cin] and image n uses w[n] (batched X * Y^T).
 * out = act(alpha * conv(x, w) + bias + residual). R,S in {1,3}; stride in {1,2}. */
int mfr_conv2d_nhwc_tf32(const float* x, int N, int H, int W, int Cin, long long ldx,
                         const float* w, int Cout, int R, int S, int stride, int pad,
                         int b_batched, const float* bias, const float* residual, long long ldr,
                         int act, float alpha, float* out, long long ldo, void* stream);

/* Same op with bf16 operands (x, w as bf16 bits; tcgen05 kind::f16, fp32 accumulate). ldx / ldw are
 * element strides with ldx*2 and ldw*2 multiples of 16 bytes (ldw = per-tap row stride of w, >= Cin).
 * out / residual are bf16 when out_bf16 != 0 (then Cout % 4 == 0), else fp32. bias is fp32.
 * bf16 output rows whose stride ldo is exactly Cout rounded up to a multiple of 8 own their padding: the
 * (at most 7) elements past Cout of every pixel may be overwritten with zeros (16-byte store units). */
int mfr_conv2d_nhwc_bf16(const void* x, int N, int H, int W, int Cin, long long ldx, const void* w,
                         long long ldw, int Cout, int R, int S, int stride, int pad, const float* bias,
                         const void* residual, long long ldr, int act, float alpha, void* out,
                         long long ldo, int out_bf16, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LoFTR matcher engine. Replaces LoFTR.forward of the reference
 * (etc/feature_matching_baselines/LoFTR/src/loftr/loftr.py:29-75) as driven by
 * LoFTR_matcher.match (etc/feature_matching_baselines/matchers.py:24-59), for B image pairs per
 * call. Images are device fp32 [B][src_h][src_w] in [0,1]; the engine frame [H][W] (multiples of 8)
 * is zero-padded on the right/bottom like the reference's ConstantPad2d (matchers.py:41-46). Weights are set tensor by tensor in the engine's own layout (BN folded,
 * conv weights [Cout][R*S][Cin]); mfr_loftr_weight_name/numel enumerate what is expected.
 * Outputs stay on the device (pointers owned by the engine), ordered by (pair, i) exactly like
 * torch.where in coarse_matching.py:194. */
typedef struct mfr_loftr mfr_loftr;
typedef struct {
  const int* m_total;    /* device int[4]: {M, 2M, 25M, 50M}, M = matches in the batch */
  const int* m_offsets;  /* device int[B+1] */
  const int* b_ids; const int* i_ids; const int* j_ids; /* device int[capacity] */
  const float* mconf;                                   /* device float[capacity] */
  const float* mkpts0_f; const float* mkpts1_f;         /* device float[capacity][2], pixels (x,y) */
  const float* mkpts1_c; const float* expec_f;          /* device float[capacity][2] / [capacity][3] */
  int capacity;
} mfr_loftr_outputs;

/* precision:
 *  0 TF32   fp32 storage everywhere, TF32 tensor-core inputs in every conv / linear / correlation GEMM (what
 *           torch + cuDNN do by default for the reference's convs on a GPU); attention, LayerNorm, softmax fp32.
 *  1 BF16   throughput mode. bf16 storage + kind::f16 MMAs (fp32 accumulate) for the ResNet-FPN maps, the token
 *           GEMM operands (q|k|v, mlp0 read a bf16 shadow of the token stream; q|k|v, attention message and MLP
 *           hidden tensors are stored in bf16), the linear attention and the coarse correlation operands. The two
 *           backbone outputs, the token residual stream, LayerNorm arithmetic, the similarity matrix S, dual
 *           softmax and fine matching stay fp32.
 *  2 FP32X3 fp32-faithful parity mode. fp32 storage; every GEMM (convs, linears, correlation) runs as a 3xTF32
 *           split product  A.B ~= Ahi.Bhi + Alo.Bhi + Ahi.Blo  (one tensor-core GEMM over 3x the K extent,
 *           ~2^-21 relative per product instead of TF32's 2^-11); stem, attention, LayerNorm in plain fp32. This
 *           is the mode whose match indices are compared with the reference's fp32 arithmetic. */
#define MFR_PRECISION_TF32 0
#define MFR_PRECISION_BF16 1
#define MFR_PRECISION_FP32X3 2
int mfr_loftr_create(int B, int H, int W, float thr, int border_rm, float temperature, int precision,
                     mfr_loftr** out);
void mfr_loftr_destroy(mfr_loftr* h);
int mfr_loftr_num_weights(const mfr_loftr* h);
const char* mfr_loftr_weight_name(const mfr_loftr* h, int i);
long long mfr_loftr_weight_numel(const mfr_loftr* h, int i);
int mfr_loftr_set_weight(mfr_loftr* h, int i, const float* host, long long numel, void* stream);
int mfr_loftr_forward(mfr_loftr* h, const float* img0, const float* img1, int src_h, int src_w,
                      void* stream);
void mfr_loftr_set_graph(mfr_loftr* h, int enabled);
int mfr_loftr_get_outputs(mfr_loftr* h, mfr_loftr_outputs* out);
/* named intermediate buffers for stage-wise parity tests ("feat_c_backbone", "feat_f", "tok", "sim", ...) */
int mfr_loftr_debug_tensor(mfr_loftr* h, const char* name, float** ptr, long long* numel);
int mfr_loftr_num_kernels(const mfr_loftr* h);  /* steps of the launch list per forward */
int mfr_loftr_num_launches(const mfr_loftr* h); /* kernel launches per forward */
/* Eager run with a CUDA-event pair around every step of the launch list (host arrays out):
 * ms[i] device time, flops[i] algorithmic FLOPs, kinds[i] 1 = tcgen05 conv/GEMM, names 32 bytes each.
 * Returns the number of steps (or a negative MFR_ERR_*). */
int mfr_loftr_profile(mfr_loftr* h, const float* img0, const float* img1, int src_h, int src_w,
                      void* stream, int max_steps, float* ms, double* flops, int* kinds, char* names);

/* ---------------------------------------------------------------------------------------------
 * SuperPoint + SuperGlue matcher engine. Replaces Matching.forward of the reference
 * (etc/feature_matching_baselines/SuperGlue/models/matching.py:56-84 = superpoint.py:145-202 x2 +
 * superglue.py:228-283) as driven by SuperGlue_matcher.match (matchers.py:93-120), B pairs per
 * call, images device fp32 [B][H][W] in [0,1]. Settings of matchers.py:65-71 are the defaults of
 * the Python wrapper (nms 4, threshold 0.005, border 4, 1024 keypoints, 20 Sinkhorn iterations,
 * match threshold 0.2). Weights in the engine's layout (mfr_spsg_weight_name enumerates them). */
typedef struct mfr_spsg mfr_spsg;
typedef struct {
  const int* counts;             /* device int[2B], pair-interleaved (2p = image0 of pair p) */
  const float* keypoints;        /* device float[2B][K][2] pixel (x,y) */
  const float* scores;           /* device float[2B][K] */
  const float* descriptors;      /* device float[2B][K][desc_ld] */
  int desc_ld;
  const int* matches0;           /* device int[B][K], -1 = unmatched (superglue.py:274) */
  const float* matching_scores0; /* device float[B][K] */
  const int* m_offsets;          /* device int[B+1] */
  const float* mkpts0; const float* mkpts1; /* device float[total][2], ordered by (pair, i) */
  int max_keypoints;
} mfr_spsg_outputs;
int mfr_spsg_create(int B, int H, int W, int max_keypoints, int nms_radius, float keypoint_threshold,
                    int remove_borders, int sinkhorn_iterations, float match_threshold, mfr_spsg** out);
/* same with the GEMM precision: MFR_PRECISION_TF32 (what mfr_spsg_create uses) or MFR_PRECISION_FP32X3 (every conv /
 * linear / attention / score GEMM as a 3xTF32 split product: the fp32-faithful parity mode) */
int mfr_spsg_create_ex(int B, int H, int W, int max_keypoints, int nms_radius, float keypoint_threshold,
                       int remove_borders, int sinkhorn_iterations, float match_threshold, int precision,
                       mfr_spsg** out);
void mfr_spsg_destroy(mfr_spsg* h);
int mfr_spsg_num_weights(const mfr_spsg* h);
const char* mfr_spsg_weight_name(const mfr_spsg* h, int i);
long long mfr_spsg_weight_numel(const mfr_spsg* h, int i);
int mfr_spsg_set_weight(mfr_spsg* h, int i, const float* host, long long numel, void* stream);
int mfr_spsg_forward(mfr_spsg* h, const float* img0, const float* img1, void* stream);
void mfr_spsg_set_graph(mfr_spsg* h, int enabled);
int mfr_spsg_get_outputs(mfr_spsg* h, mfr_spsg_outputs* out);
int mfr_spsg_debug_tensor(mfr_spsg* h, const char* name, float** ptr, long long* numel);
int mfr_spsg_num_launches(const mfr_spsg* h);
/* SuperGlue alone (superglue.py:228-283) on caller-supplied SuperPoint outputs, HOST arrays in the
 * engine's layout: counts int[2B], keypoints float[2B][K][2], scores float[2B][K], descriptors
 * float[2B][K][256] (unused slots zero). Stage-boundary entry point. */
int mfr_spsg_forward_superglue(mfr_spsg* h, const int* counts, const float* keypoints, const float* scores,
                               const float* descriptors, void* stream);

/* Coarse matching on a materialised matrix (stage-boundary entry point of the LoFTR matcher):
 * CoarseMatching.forward dual-softmax + get_coarse_match of the reference
 * (LoFTR/src/loftr/utils/coarse_matching.py:109-119,175-196,242-259).
 *   sim: device float[B][L][L] similarity (already scaled by 1/(C*temperature)), or, when
 *   input_is_conf != 0, the confidence matrix itself (then the selection is bit-exact w.r.t. the
 *   reference's float-equality mutual-NN test on the same matrix).
 * Outputs (device, caller-owned): count int[1], offsets int[B+1], b/i/j_ids int[B*L],
 * mconf float[B*L], mkpts0_c/mkpts1_c float[B*L][2] = (idx % w_c, idx / w_c) * scale. */
int mfr_coarse_match(const float* sim, int B, int h_c, int w_c, float thr, int border_rm, float scale,
                     int input_is_conf, int* count, int* offsets, int* b_ids, int* i_ids, int* j_ids,
                     float* mconf, float* mkpts0_c, float* mkpts1_c, void* stream);

/* Exact 2-nearest-neighbour search (L2) of every query descriptor among the train descriptors:
 * the matcher of the reference's SIFT plugin (lib/models/matching/feature_matching.py:87-99),
 * knnMatch(k=2), as an exact brute-force search (the reference's FLANN kd-tree is approximate and
 * non-deterministic). query float[nq][dim], train float[nt][dim], dim % 4 == 0, dim <= 256 (device).
 * Outputs (device): idx1/idx2 int[nq] (-1 if absent), dist1/dist2 float[nq] (Euclidean). Ties go
 * to the lower index. */
int mfr_knn2_l2(const float* query, int nq, const float* train, int nt, int dim, int* idx1,
                float* dist1, int* idx2, float* dist2, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Pose back end (batched-hypothesis RANSAC). Replaces lib/models/matching/pose_solver.py.
 * Correspondences of P pairs are concatenated: kpts0/kpts1 device float[total][2] in pixels (x,y),
 * offsets device int[P+1]; K0/K1 device float[P][9] row-major intrinsics; depth maps device
 * float[P][H][W] metres (0 = invalid). Outputs: R float[P][9], t float[P][3], inliers int[P],
 * mask unsigned char[total]. Failure of a pair = NaN-filled R,t and 0 inliers
 * (pose_solver.py:30-33,50-51,146-150,189,198,231-233).
 * A solver handle owns the scratch memory; correspondences per pair are capped at 8192.       */
typedef struct mfr_pose_solver mfr_pose_solver;
int mfr_pose_solver_create(int max_pairs, int max_pts_per_pair, int max_hypotheses,
                           mfr_pose_solver** out);
void mfr_pose_solver_destroy(mfr_pose_solver* s);

/* EssentialMatrixSolver.estimate_pose (pose_solver.py:29-61): cv.findEssentialMat(USAC_MAGSAC) +
 * cv.recoverPose. `num_hyp` 5-point hypotheses per pair (MSAC-scored). lo_enabled is a bit set:
 *   bit 0 (1): local optimisation of the 4 best block winners (LM on a Tukey M-estimator of the Sampson distance,
 *              cut-off 3x -> 1x the threshold) + noise-adaptive tightening;
 *   bit 1 (2): final truncated least-squares refit at the threshold (LM on the MSAC cost): the result is the Sampson
 *              least-squares optimum of its OWN inlier set — the point the reference's output polished by LM on its
 *              own mask sits at (SURVEY §7 hard part 1(b)); needs bit 0.
 * t has unit norm; max_pts_per_pair <= 8192 (more correspondences of a pair are ignored). The final (R, t) is also
 * kept in fp64 inside the handle for mfr_scale_from_depth. */
int mfr_emat_ransac(mfr_pose_solver* s, const float* kpts0, const float* kpts1, const int* offsets,
                    int P, const float* K0, const float* K1, float pix_threshold, int num_hyp,
                    unsigned long long seed, int lo_enabled, float* R, float* t, int* n_inliers,
                    unsigned char* mask, void* stream);

/* EssentialMatrixMetricSolver.estimate_pose (pose_solver.py:132-172), the part after the
 * scale-free solve: depth lookup at int-truncated inlier keypoints, back-projection, per-point
 * scale, exhaustive consensus with |s - s_hyp| < scale_threshold. Rd/td: device double[P][9]/[P][3]
 * pose to use, or NULL to take the pose left in the handle by mfr_emat_ransac. */
int mfr_scale_from_depth(mfr_pose_solver* s, const float* kpts0, const float* kpts1,
                         const int* offsets, int P, const float* K0, const float* K1,
                         const unsigned char* mask, const double* Rd, const double* td,
                         const float* depth0, const float* depth1, int H, int W,
                         float scale_threshold, float* R_out, float* t_metric, int* inliers,
                         void* stream);

/* PnPSolver.estimate_pose (pose_solver.py:184-235): lift kpts0 through depth0 (valid iff depth >
 * depth0.min()), P3P RANSAC on (xyz0, kpts1) with reprojection threshold in pixels, LM refinement
 * on the inliers when there are >= 6, reject |t| > 1000. */
int mfr_pnp_ransac(mfr_pose_solver* s, const float* kpts0, const float* kpts1, const int* offsets,
                   int P, const float* K0, const float* K1, const float* depth0, int H, int W,
                   int num_hyp, float reproj_threshold, unsigned long long seed, float* R, float* t,
                   int* n_inliers, unsigned char* mask, void* stream);

/* ProcrustesSolver.estimate_pose with REFINE = False (pose_solver.py:247-320): lift both keypoint sets
 * through their depth maps (valid iff depth > depth.min() on both sides), RANSAC over 3-point Kabsch
 * alignments with inlier distance max_corr_dist (metres), refit on the inliers. The reference calls
 * Open3D's registration_ransac_based_on_correspondence for this step; REFINE = True continues with mfr_icp_refine. */
int mfr_procrustes_ransac(mfr_pose_solver* s, const float* kpts0, const float* kpts1, const int* offsets,
                          int P, const float* K0, const float* K1, const float* depth0,
                          const float* depth1, int H, int W, int num_hyp, float max_corr_dist,
                          unsigned long long seed, float* R, float* t, int* n_inliers, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Host test hooks: the minimal solvers used inside the RANSAC kernels are __host__ __device__
 * functions; these run the very same code on the CPU so that it can be unit-tested without a GPU.
 * They are NOT a CPU fallback: no product entry point calls them.
 *   five_point : x0, x1 double[5][2] normalised points, E_out double[10][9]; returns #solutions
 *   p3p        : X double[3][3] points, f double[3][3] unit bearings, R_out [4][9], t_out [4][3]
 *   poly roots : coef[0..deg] ascending, deg <= 10; returns #real roots (ascending)          */
int mfr_host_five_point(const double* x0, const double* x1, double* E_out);
int mfr_host_p3p(const double* X, const double* f, double* R_out, double* t_out);
int mfr_host_poly_real_roots(const double* coef, int deg, double* roots);
int mfr_host_decompose_essential(const double* E, double* R1, double* R2, double* t);

/* Debug probes (microbenchmarks used by scripts/, not part of the product path): microseconds per
 * launch of an empty / TMEM-allocating kernel with the GEMM's launch shape, and of a prebuilt
 * conv/GEMM op launched back to back. */
float mfr_debug_probe_overhead(int mode, int smem_bytes, int grid, int iters);
/* tensor-pipe probe: ns per tcgen05.mma (128 x BN x 32 B of K, operands resident in shared memory) when one
 * thread issues n_per_commit MMAs per tcgen05.commit, optionally waiting for each commit */
float mfr_debug_mma_probe(int n_per_commit, int BN, int bf16, int wait_each, int iters, int grid, int two_commits);
/* flags: bit 0 = x/w are bf16, bit 1 = out/residual are bf16, bit 2 = channel strides padded to multiples of 8 elements;
 * residual may be NULL */
float mfr_debug_time_conv(const float* x, int N, int H, int W, int Cin, const float* w, int Cout, int R,
                          int stride, float* out, int iters, int flags, const float* residual);

/* Stage entry for tests: LoFTR's linear attention (LoFTR/src/loftr/loftr_module/linear_attention.py:31-45) on
 * already feature-mapped q' = elu(q)+1, k' and values. windows = 0: q [n][L][256], k / v [n][S][256], 8 heads of 32
 * (coarse level); windows = 1: q, k, v [n][25][128], 8 heads of 16, n items (fine level, L and S ignored).
 * bf16 = 1: all four tensors are bf16 (the tensor-core kernels), else fp32. bf16 = 2: bf16 tensors with RAW q and k (the
 * projections as the GEMM leaves them): the kernels apply elu(x) + 1 themselves. Device pointers, dense rows. */
int mfr_linear_attention(const void* q, const void* k, const void* v, void* msg, int n, int L, int S, int windows,
                         int bf16, void* stream);

/* Stage entry for tests: the backbone stem, conv 7x7 stride 2 pad 3, 1 -> 128 channels, folded-BN bias, ReLU
 * (LoFTR/src/loftr/backbone/resnet_fpn.py:52-62,102). img fp32 [NI][H][W]; w fp32 [128][49]; out [NI][H/2][W/2][128], bf16 when
 * out_bf16 != 0 (the tcgen05 kernel; MFR_STEM_TC=0 selects the mma.sync one), else fp32; exact_fp32: plain FMA kernel. */
int mfr_stem_conv7x7(const float* img, const float* w, const float* bias, void* out, int NI, int H, int W, int out_bf16,
                     int exact_fp32, void* stream);

/* Stage entry for tests: the fine level's q|k|v projection + window attention as ONE kernel (bf16 mode; the projection is
 * LoFTREncoderLayer's q_proj / k_proj / v_proj, transformer.py:21-23,41-46, the attention linear_attention.py:31-45 on 5x5
 * windows). xq / xs: bf16 token rows [n_items * 25][>= 128] of the query / source side with element strides ldq / lds (the same
 * pointer = self attention); w_qkv: bf16 [384][128] = rows of Wq, Wk, Wv; msg: bf16 [n_items * 25][ldm]. Device pointers. */
int mfr_fine_attention_fused(const void* xq, long long ldq, const void* xs, long long lds, const void* w_qkv, void* msg,
                             long long ldm, int n_items, void* stream);

/* ProcrustesSolver's ICP refinement (PROCRUSTES.REFINE, pose_solver.py:289-315: Open3D registration_icp, point-to-point,
 * correspondence radius max_corr_dist, criteria relative_fitness / relative_rmse / max_iterations = 1e-4 / 1e-4 / 30 in the
 * reference) of the poses (R_in, t_in) [P][9] / [P][3] over the FULL clouds of both depth maps (every pixel with depth > 0),
 * device float [P][H][W]. Nearest neighbours are exact within the radius (bounded window search in the organised target
 * cloud). n_inliers = int(fitness * valid target points), the reference's return value (pose_solver.py:319). */
int mfr_icp_refine(mfr_pose_solver* s, const float* depth0, const float* depth1, int H, int W, int P, const float* K0,
                   const float* K1, const float* R_in, const float* t_in, float max_corr_dist, int max_iterations,
                   float relative_fitness, float relative_rmse, float* R, float* t, int* n_inliers, void* stream);

/* Input path on the GPU. mfr_jpeg_info / mfr_jpeg_decode: nvJPEG decode of a JPEG held in HOST
 * memory into a DEVICE uint8 image — rgb = 0: the luma plane [h][w] (cv2.imread(IMREAD_GRAYSCALE) of a JPEG,
 * SuperGlue/models/utils.py:264), rgb = 1: interleaved RGB [h][w][3] (lib/datasets/utils.py:67 imread).
 * mfr_resize_linear: cv2.resize(INTER_LINEAR) + scale on the device, dst fp32 [dst_h][dst_w][channels] (chw = 0) or
 * [channels][dst_h][dst_w] (chw = 1) = resized * out_scale. mode 0: float source, float arithmetic; mode 1: uint8 source
 * converted to float first (read_image with resize_float = True, utils.py:272-273: the FMB matcher plugins);
 * mode 2: uint8 source, OpenCV's 11-bit fixed-point uint8 path, rounded to uint8 before the scale (utils.py:275,
 * lib/datasets/utils.py:68). src_row_stride in elements. */
int mfr_jpeg_info(const unsigned char* data, long long length, int* width, int* height, int* components);
int mfr_jpeg_decode(const unsigned char* data, long long length, int rgb, unsigned char* dst, int width, int height, void* stream);
int mfr_resize_linear(const void* src, int src_is_u8, int src_w, int src_h, long long src_row_stride, int channels, float* dst,
                      int dst_w, int dst_h, int mode, float out_scale, int chw, void* stream);

/* Leaderboard metrics of the Map-free benchmark for n poses (benchmark/metrics.py:40-67, benchmark/utils.py:95-129 sine
 * variant, benchmark/reprojection.py:7-86): device double arrays q_gt/q_est [n][4] (w, x, y, z), t_gt/t_est [n][3],
 * K [n][9]; image size W x H. Out: trans_err [n] (m), rot_err [n] (degrees), reproj_err [n] (VCRE, pixels). */
int mfr_pose_metrics(const double* q_gt, const double* t_gt, const double* q_est, const double* t_est, const double* K, int W,
                     int H, int n, double* trans_err, double* rot_err, double* reproj_err, void* stream);

/* Stage entry: the fused tail of one LoFTREncoderLayer (loftr_module/transformer.py:52-58),
 *   x += norm2(mlp2(relu(mlp0([x | norm1(merge(msg))])))),  bf16 tensor-core operands, fp32 accumulate / LayerNorm / residual,
 * for n_blocks row blocks of rows_cap token rows (live rows per block: *dyn_rows when non-null). Device pointers:
 * msg bf16 [rows][C]; xb bf16 [rows][2C] (columns [0, C) = shadow of x: read, then rewritten with the new x);
 * x fp32 [rows][2C] (columns [0, C) in / out); w_merge [C][C], w_mlp0 [2C][2C], w_mlp2 [C][2C] bf16 row-major
 * ([out][in], the reference's nn.Linear layout); g1 / b1 / g2 / b2 fp32 [C]. C = 256 or 128. `iters` launches back to
 * back (timing); note that every launch applies the residual update again. */
int mfr_encoder_tail(const void* msg, void* xb, float* x, const void* w_merge, const void* w_mlp0, const void* w_mlp2,
                     const float* g1, const float* b1, const float* g2, const float* b2, int C, int n_blocks,
                     long long rows_cap, const int* dyn_rows, int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MFR_B200_H_ */
