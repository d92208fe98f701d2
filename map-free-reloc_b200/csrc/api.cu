// C ABI of libmfr_b200.so (declared in include/mfr_b200.h).
#include "../../include/mfr_b200.h"

#include "common.cuh"
#include "conv_gemm.cuh"

extern "C" {

int mfr_version(void) { return 100; }

int mfr_device_sm_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return 0;
  return mfr::conv_gemm_num_sms();
}

int mfr_conv2d_nhwc_tf32(const float* x, int N, int H, int W, int Cin, long long ldx,
                         const float* w, int Cout, int R, int S, int stride, int pad,
                         int b_batched, const float* bias, const float* residual, long long ldr,
                         int act, float alpha, float* out, long long ldo, void* stream) {
  mfr::ConvGemmDesc d{};
  d.x = x; d.N = N; d.H = H; d.W = W; d.Cin = Cin; d.ldx = ldx;
  d.w = w; d.Cout = Cout; d.R = R; d.S = S; d.stride = stride; d.pad = pad;
  d.b_batched = b_batched; d.bias = bias; d.residual = residual; d.ldr = ldr;
  d.act = act; d.alpha = alpha; d.out = out; d.ldo = ldo;
  mfr::ConvGemmParams p;
  int rc = mfr::conv_gemm_build(d, &p);
  if (rc != MFR_OK) return rc;
  return mfr::conv_gemm_launch(p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
