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

#ifdef __cplusplus
}
#endif
#endif /* MFR_B200_H_ */
