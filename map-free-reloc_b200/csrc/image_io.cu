// Input path of the matchers on the GPU: JPEG decode (nvJPEG, luma plane = what
// cv2.imread(IMREAD_GRAYSCALE) returns for a JPEG) + the resize / normalise step of the reference's loaders
//   read_image      etc/feature_matching_baselines/SuperGlue/models/utils.py:263-285   gray, cv2.resize, / 255
//   read_color_image lib/datasets/utils.py:58-75                                        RGB uint8 cv2.resize, / 255, CHW
// cv2.resize(INTER_LINEAR) is restated for both of its code paths: float32 images (bilinear weights in fp32, horizontal
// pass first) and uint8 images (11-bit fixed-point weights, OpenCV's two-stage rounding), so that the resized pixels are
// the ones the reference feeds its networks. 16-bit depth PNGs (read_depth_image) stay on the host: inflate is not GPU work.
#include <nvjpeg.h>

#include <mutex>

#include "common.cuh"

namespace mfr {

namespace {

// source coordinate of destination pixel d (cv2: fx = (dx + 0.5) * scale - 0.5, clamped like resizeGeneric's tables)
__device__ __forceinline__ void src_coord(int d, double scale, int ssize, int& s0, float& f) {
  float fx = static_cast<float>((d + 0.5) * scale - 0.5);
  int sx = static_cast<int>(floorf(fx));
  fx -= sx;
  if (sx < 0) { fx = 0.f; sx = 0; }
  if (sx >= ssize - 1) { fx = 0.f; sx = ssize - 1; }
  s0 = sx; f = fx;
}

// float path: dst = ((1-fy) * ((1-fx) S00 + fx S01) + fy * ((1-fx) S10 + fx S11)) * out_scale, SRC = float or uint8 (converted first)
template <typename SRC>
__global__ void resize_linear_f32_kernel(const SRC* __restrict__ src, int sw, int sh, long long src_stride, int channels,
                                         float* __restrict__ dst, int dw, int dh, double sx, double sy, float out_scale, int chw) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dw) return;
  int x0, y0; float fx, fy;
  src_coord(x, sx, sw, x0, fx);
  src_coord(y, sy, sh, y0, fy);
  const int x1 = min(x0 + 1, sw - 1), y1 = min(y0 + 1, sh - 1);
  for (int c = 0; c < channels; ++c) {
    const float s00 = static_cast<float>(src[y0 * src_stride + x0 * channels + c]), s01 = static_cast<float>(src[y0 * src_stride + x1 * channels + c]);
    const float s10 = static_cast<float>(src[y1 * src_stride + x0 * channels + c]), s11 = static_cast<float>(src[y1 * src_stride + x1 * channels + c]);
    const float r0 = __fadd_rn(__fmul_rn(s00, 1.f - fx), __fmul_rn(s01, fx));      // HResizeLinear: S[x] * a0 + S[x+1] * a1
    const float r1 = __fadd_rn(__fmul_rn(s10, 1.f - fx), __fmul_rn(s11, fx));
    const float v = __fadd_rn(__fmul_rn(r0, 1.f - fy), __fmul_rn(r1, fy));        // VResizeLinear: S0 * b0 + S1 * b1
    const size_t o = chw ? (static_cast<size_t>(c) * dh + y) * dw + x : (static_cast<size_t>(y) * dw + x) * channels + c;
    dst[o] = v * out_scale;
  }
}

// uint8 path (cv2.resize on uint8): INTER_RESIZE_COEF_BITS = 11 fixed-point weights, rows kept as int (x 2048), vertical
// pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2, saturated to uint8; then float * out_scale
__global__ void resize_linear_u8_kernel(const unsigned char* __restrict__ src, int sw, int sh, long long src_stride, int channels,
                                        float* __restrict__ dst, int dw, int dh, double sx, double sy, float out_scale, int chw) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dw) return;
  int x0, y0; float fx, fy;
  src_coord(x, sx, sw, x0, fx);
  src_coord(y, sy, sh, y0, fy);
  const int x1 = min(x0 + 1, sw - 1), y1 = min(y0 + 1, sh - 1);
  auto coef = [](float v) { return static_cast<int>(max(-32768, min(32767, __float2int_rn(v * 2048.f)))); };   // saturate_cast<short>(cvRound)
  const int a0 = coef(1.f - fx), a1 = coef(fx), b0 = coef(1.f - fy), b1 = coef(fy);
  for (int c = 0; c < channels; ++c) {
    const int r0 = src[y0 * src_stride + x0 * channels + c] * a0 + src[y0 * src_stride + x1 * channels + c] * a1;
    const int r1 = src[y1 * src_stride + x0 * channels + c] * a0 + src[y1 * src_stride + x1 * channels + c] * a1;
    int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
    v = max(0, min(255, v));
    const size_t o = chw ? (static_cast<size_t>(c) * dh + y) * dw + x : (static_cast<size_t>(y) * dw + x) * channels + c;
    dst[o] = static_cast<float>(v) * out_scale;
  }
}

struct Jpeg {
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  bool ok = false;
};
Jpeg& jpeg() {
  static Jpeg j;
  static std::once_flag once;
  std::call_once(once, [] {
    if (nvjpegCreateSimple(&j.handle) != NVJPEG_STATUS_SUCCESS) return;
    if (nvjpegJpegStateCreate(j.handle, &j.state) != NVJPEG_STATUS_SUCCESS) return;
    j.ok = true;
  });
  return j;
}
std::mutex g_jpeg_mutex;

}  // namespace

// mode 0: float path on a float32 source; 1: float path on a uint8 source (image.astype('float32') then resize: read_image with
// resize_float=True); 2: uint8 fixed-point path (cv2.resize on uint8, then float). dst = resized * out_scale.
int resize_linear(const void* src, int src_is_u8, int sw, int sh, long long src_stride, int channels, float* dst, int dw, int dh,
                  int mode, float out_scale, int chw, cudaStream_t st) {
  if (!src || !dst || sw < 1 || sh < 1 || dw < 1 || dh < 1 || channels < 1 || channels > 4) return MFR_ERR_ARG;
  const double sx = static_cast<double>(sw) / dw, sy = static_cast<double>(sh) / dh;
  const dim3 grid((dw + 127) / 128, dh), block(128);
  if (mode == 2) {
    if (!src_is_u8) return MFR_ERR_ARG;
    resize_linear_u8_kernel<<<grid, block, 0, st>>>(static_cast<const unsigned char*>(src), sw, sh, src_stride, channels, dst, dw, dh, sx, sy, out_scale, chw);
  } else if (src_is_u8) {
    resize_linear_f32_kernel<unsigned char><<<grid, block, 0, st>>>(static_cast<const unsigned char*>(src), sw, sh, src_stride, channels, dst, dw, dh, sx, sy, out_scale, chw);
  } else {
    resize_linear_f32_kernel<float><<<grid, block, 0, st>>>(static_cast<const float*>(src), sw, sh, src_stride, channels, dst, dw, dh, sx, sy, out_scale, chw);
  }
  MFR_CUDA_CHECK(cudaGetLastError());
  return MFR_OK;
}

int jpeg_info(const unsigned char* data, size_t length, int* width, int* height, int* components) {
  Jpeg& j = jpeg();
  if (!j.ok) return MFR_ERR_UNSUPPORTED;
  int nc = 0;
  nvjpegChromaSubsampling_t ss;
  int ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
  std::lock_guard<std::mutex> lk(g_jpeg_mutex);
  if (nvjpegGetImageInfo(j.handle, data, length, &nc, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS) return MFR_ERR_ARG;
  *width = ws[0]; *height = hs[0]; *components = nc;
  return MFR_OK;
}

// Decodes a JPEG (host bytes) into a device uint8 image: gray [h][w] (the luma plane) or interleaved RGB [h][w][3].
int jpeg_decode(const unsigned char* data, size_t length, int rgb, unsigned char* dst, int width, int height, cudaStream_t st) {
  Jpeg& j = jpeg();
  if (!j.ok) return MFR_ERR_UNSUPPORTED;
  nvjpegImage_t img{};
  img.channel[0] = dst;
  img.pitch[0] = static_cast<size_t>(width) * (rgb ? 3 : 1);
  std::lock_guard<std::mutex> lk(g_jpeg_mutex);
  const nvjpegStatus_t s = nvjpegDecode(j.handle, j.state, data, length, rgb ? NVJPEG_OUTPUT_RGBI : NVJPEG_OUTPUT_Y, &img, st);
  if (s != NVJPEG_STATUS_SUCCESS) {
    fprintf(stderr, "[mfr_b200] nvjpegDecode failed: %d\n", static_cast<int>(s));
    return MFR_ERR_CUDA;
  }
  return MFR_OK;
}

}  // namespace mfr
