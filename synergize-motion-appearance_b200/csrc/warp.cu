// Stage 2: bilinear warp of encoded source features fused with the on-the-fly align_corners=True
// resize of the 64x64 deformation and occlusion maps and the occlusion multiply; plus the generic
// NHWC bilinear resize.  Pure HBM/L2-bound gathers: NHWC makes every corner fetch a contiguous
// 16-byte-vector run over channels, one output pixel per group of C/4 lanes.
#include "sma_common.cuh"

namespace {

struct Bil { int i0, i1; float w1; };
// source coordinate for output index o under align_corners=True, as ATen computes it
// (area_pixel_compute_source_index: scale = (in-1)/(out-1), src = scale*o, i0 = floor, lambda = src-i0)
__device__ __forceinline__ Bil bil_ac(int o, int n_in, int n_out) {
  Bil r;
  float scale = n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f;
  float src = scale * o;
  int i0 = (int)src;                   // src >= 0
  if (i0 > n_in - 1) i0 = n_in - 1;
  r.i0 = i0; r.i1 = i0 + (i0 < n_in - 1 ? 1 : 0); r.w1 = src - (float)i0;
  return r;
}

__device__ __forceinline__ float4 f4_fma(float w, float4 a, float4 acc) {
  acc.x = fmaf(w, a.x, acc.x); acc.y = fmaf(w, a.y, acc.y); acc.z = fmaf(w, a.z, acc.z); acc.w = fmaf(w, a.w, acc.w);
  return acc;
}

// one thread per (pixel, channel-quad)
__global__ void warp_occlude_kernel(const float* __restrict__ feat, long long fbs, int B, int H, int W, int C,
                                    const float* __restrict__ flow, const float* __restrict__ occ, int hf, int wf,
                                    float* __restrict__ out, long long total4) {
  const int C4 = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4); long long pp = i / C4; int x = (int)(pp % W); long long t = pp / W; int y = (int)(t % H); int b = (int)(t / H);
    // resized flow / occlusion at (y,x)
    float gx, gy, oc = 1.f;
    if (hf == H && wf == W) {
      const float* f = flow + (((long long)b * hf + y) * wf + x) * 2; gx = __ldg(f); gy = __ldg(f + 1);
      if (occ) oc = __ldg(occ + ((long long)b * hf + y) * wf + x);
    } else {
      Bil by = bil_ac(y, hf, H), bx = bil_ac(x, wf, W);
      const float* f = flow + (long long)b * hf * wf * 2;
      float2 f00 = __ldg(reinterpret_cast<const float2*>(f + ((long long)by.i0 * wf + bx.i0) * 2));
      float2 f01 = __ldg(reinterpret_cast<const float2*>(f + ((long long)by.i0 * wf + bx.i1) * 2));
      float2 f10 = __ldg(reinterpret_cast<const float2*>(f + ((long long)by.i1 * wf + bx.i0) * 2));
      float2 f11 = __ldg(reinterpret_cast<const float2*>(f + ((long long)by.i1 * wf + bx.i1) * 2));
      float w0y = 1.f - by.w1, w0x = 1.f - bx.w1;
      gx = w0y * (w0x * f00.x + bx.w1 * f01.x) + by.w1 * (w0x * f10.x + bx.w1 * f11.x);
      gy = w0y * (w0x * f00.y + bx.w1 * f01.y) + by.w1 * (w0x * f10.y + bx.w1 * f11.y);
      if (occ) {
        const float* o = occ + (long long)b * hf * wf;
        float o00 = __ldg(o + by.i0 * wf + bx.i0), o01 = __ldg(o + by.i0 * wf + bx.i1);
        float o10 = __ldg(o + by.i1 * wf + bx.i0), o11 = __ldg(o + by.i1 * wf + bx.i1);
        oc = w0y * (w0x * o00 + bx.w1 * o01) + by.w1 * (w0x * o10 + bx.w1 * o11);
      }
    }
    // grid_sample, align_corners=True: ix = (gx+1)/2*(W-1)
    float ix = (gx + 1.f) * 0.5f * (float)(W - 1), iy = (gy + 1.f) * 0.5f * (float)(H - 1);
    float fx = floorf(ix), fy = floorf(iy);
    int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const float* fb = feat + (long long)b * fbs + c4 * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)x1 < (unsigned)W, vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)y1 < (unsigned)H;
    // same accumulation order as ATen's grid_sampler_2d (nw, ne, sw, se)
    if (vy0 && vx0) acc = f4_fma(wy0 * wx0, __ldg(reinterpret_cast<const float4*>(fb + ((long long)y0 * W + x0) * C)), acc);
    if (vy0 && vx1) acc = f4_fma(wy0 * wx1, __ldg(reinterpret_cast<const float4*>(fb + ((long long)y0 * W + x1) * C)), acc);
    if (vy1 && vx0) acc = f4_fma(wy1 * wx0, __ldg(reinterpret_cast<const float4*>(fb + ((long long)y1 * W + x0) * C)), acc);
    if (vy1 && vx1) acc = f4_fma(wy1 * wx1, __ldg(reinterpret_cast<const float4*>(fb + ((long long)y1 * W + x1) * C)), acc);
    if (occ) { acc.x *= oc; acc.y *= oc; acc.z *= oc; acc.w *= oc; }
    *reinterpret_cast<float4*>(out + (pp * C) + c4 * 4) = acc;
  }
}

__global__ void resize_ac_kernel(const float* __restrict__ x, int Hi, int Wi, int C, long long ibs, int ild,
                                 float* __restrict__ y, int Ho, int Wo, long long obs, int old_, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long pp = i / C; int ox = (int)(pp % Wo); long long t = pp / Wo; int oy = (int)(t % Ho); int b = (int)(t / Ho);
    Bil by = bil_ac(oy, Hi, Ho), bx = bil_ac(ox, Wi, Wo);
    const float* xb = x + (long long)b * ibs + c;
    float v00 = __ldg(xb + ((long long)by.i0 * Wi + bx.i0) * ild), v01 = __ldg(xb + ((long long)by.i0 * Wi + bx.i1) * ild);
    float v10 = __ldg(xb + ((long long)by.i1 * Wi + bx.i0) * ild), v11 = __ldg(xb + ((long long)by.i1 * Wi + bx.i1) * ild);
    float w0y = 1.f - by.w1, w0x = 1.f - bx.w1;
    y[(long long)b * obs + ((long long)oy * Wo + ox) * old_ + c] = w0y * (w0x * v00 + bx.w1 * v01) + by.w1 * (w0x * v10 + bx.w1 * v11);
  }
}

}  // namespace

extern "C" int sma_warp_occlude_fwd(const float* feat, int64_t fbs, int B, int H, int W, int C, const float* flow, const float* occ,
                                    int hf, int wf, float* out, sma_stream_t stream) {
  if (!feat || !flow || !out || B <= 0 || H <= 1 || W <= 1 || C <= 0 || hf <= 1 || wf <= 1) return SMA_ERR_BAD_ARG;
  if ((C & 3) || (fbs & 3) || ((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(out)) & 15) ||
      (reinterpret_cast<uintptr_t>(flow) & 7))
    return SMA_ERR_UNSUPPORTED;
  long long total4 = (long long)B * H * W * (C >> 2);
  long long blocks = (total4 + 255) / 256; if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  warp_occlude_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(feat, fbs, B, H, W, C, flow, occ, hf, wf, out, total4);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}

extern "C" int sma_resize_bilinear_ac(const float* x, int B, int Hi, int Wi, int C, int64_t ibs, int ild, float* y, int Ho, int Wo,
                                      int64_t obs, int old_, sma_stream_t stream) {
  if (!x || !y || B <= 0 || Hi <= 0 || Wi <= 0 || C <= 0 || Ho <= 0 || Wo <= 0) return SMA_ERR_BAD_ARG;
  long long total = (long long)B * Ho * Wo * C;
  long long blocks = (total + 255) / 256; if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  resize_ac_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, Hi, Wi, C, ibs, ild, y, Ho, Wo, obs, old_, total);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}
