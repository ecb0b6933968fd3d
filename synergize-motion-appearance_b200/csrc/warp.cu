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

struct F8 { float4 a, b; };
__device__ __forceinline__ F8 ldg256(const float* p) {
  F8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y),
               "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
  return r;
}

// one thread per (pixel, VEC channels); VEC = 8 uses 256-bit loads / stores (whole 32-byte sectors per lane) and halves the per-pixel
// flow / occlusion resize work that every thread of a pixel repeats
template <int VEC>
__global__ void warp_occlude_kernel(const float* __restrict__ feat, long long fbs, int B, int H, int W, int C,
                                    const float* __restrict__ flow, const float* __restrict__ occ, int hf, int wf,
                                    float* __restrict__ out, long long total4, int Hg, int Wg) {
  const int C4 = C / VEC;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4); long long pp = i / C4; int x, y, b;
    if (Hg) {      // gathered output (B, 2 Hg, 2 Wg, C): only the pixels a following bilinear down-sampling to Hg x Wg reads (its four neighbours per sample)
      const int qx = (int)(pp % (2 * Wg)); const long long t = pp / (2 * Wg); const int qy = (int)(t % (2 * Hg)); b = (int)(t / (2 * Hg));
      const Bil sy = bil_ac(qy >> 1, H, Hg), sx = bil_ac(qx >> 1, W, Wg);
      y = (qy & 1) ? sy.i1 : sy.i0; x = (qx & 1) ? sx.i1 : sx.i0;
    } else { x = (int)(pp % W); const long long t = pp / W; y = (int)(t % H); b = (int)(t / H); }
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
    const float* fb = feat + (long long)b * fbs + c4 * VEC;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)x1 < (unsigned)W, vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)y1 < (unsigned)H;
    // same accumulation order as ATen's grid_sampler_2d (nw, ne, sw, se)
    if (VEC == 8) {
      float4 acc2 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vy0 && vx0) { F8 v = ldg256(fb + ((long long)y0 * W + x0) * C); acc = f4_fma(wy0 * wx0, v.a, acc); acc2 = f4_fma(wy0 * wx0, v.b, acc2); }
      if (vy0 && vx1) { F8 v = ldg256(fb + ((long long)y0 * W + x1) * C); acc = f4_fma(wy0 * wx1, v.a, acc); acc2 = f4_fma(wy0 * wx1, v.b, acc2); }
      if (vy1 && vx0) { F8 v = ldg256(fb + ((long long)y1 * W + x0) * C); acc = f4_fma(wy1 * wx0, v.a, acc); acc2 = f4_fma(wy1 * wx0, v.b, acc2); }
      if (vy1 && vx1) { F8 v = ldg256(fb + ((long long)y1 * W + x1) * C); acc = f4_fma(wy1 * wx1, v.a, acc); acc2 = f4_fma(wy1 * wx1, v.b, acc2); }
      if (occ) { acc.x *= oc; acc.y *= oc; acc.z *= oc; acc.w *= oc; acc2.x *= oc; acc2.y *= oc; acc2.z *= oc; acc2.w *= oc; }
      asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out + (pp * C) + c4 * 8), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w),
                   "f"(acc2.x), "f"(acc2.y), "f"(acc2.z), "f"(acc2.w)
                   : "memory");
    } else {
      if (vy0 && vx0) acc = f4_fma(wy0 * wx0, __ldg(reinterpret_cast<const float4*>(fb + ((long long)y0 * W + x0) * C)), acc);
      if (vy0 && vx1) acc = f4_fma(wy0 * wx1, __ldg(reinterpret_cast<const float4*>(fb + ((long long)y0 * W + x1) * C)), acc);
      if (vy1 && vx0) acc = f4_fma(wy1 * wx0, __ldg(reinterpret_cast<const float4*>(fb + ((long long)y1 * W + x0) * C)), acc);
      if (vy1 && vx1) acc = f4_fma(wy1 * wx1, __ldg(reinterpret_cast<const float4*>(fb + ((long long)y1 * W + x1) * C)), acc);
      if (occ) { acc.x *= oc; acc.y *= oc; acc.z *= oc; acc.w *= oc; }
      *reinterpret_cast<float4*>(out + (pp * C) + c4 * 4) = acc;
    }
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

// 4 channels per thread (128-bit accesses)
__global__ void resize_ac4_kernel(const float* __restrict__ x, int Hi, int Wi, int C4, long long ibs, int ild,
                                  float* __restrict__ y, int Ho, int Wo, long long obs, int old_, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C4) * 4; long long pp = i / C4; int ox = (int)(pp % Wo); long long t = pp / Wo; int oy = (int)(t % Ho); int b = (int)(t / Ho);
    Bil by = bil_ac(oy, Hi, Ho), bx = bil_ac(ox, Wi, Wo);
    const float* xb = x + (long long)b * ibs + c;
    const float4 v00 = __ldg(reinterpret_cast<const float4*>(xb + ((long long)by.i0 * Wi + bx.i0) * ild)), v01 = __ldg(reinterpret_cast<const float4*>(xb + ((long long)by.i0 * Wi + bx.i1) * ild));
    const float4 v10 = __ldg(reinterpret_cast<const float4*>(xb + ((long long)by.i1 * Wi + bx.i0) * ild)), v11 = __ldg(reinterpret_cast<const float4*>(xb + ((long long)by.i1 * Wi + bx.i1) * ild));
    const float w0y = 1.f - by.w1, w0x = 1.f - bx.w1;
    float4 o;
    o.x = w0y * (w0x * v00.x + bx.w1 * v01.x) + by.w1 * (w0x * v10.x + bx.w1 * v11.x);
    o.y = w0y * (w0x * v00.y + bx.w1 * v01.y) + by.w1 * (w0x * v10.y + bx.w1 * v11.y);
    o.z = w0y * (w0x * v00.z + bx.w1 * v01.z) + by.w1 * (w0x * v10.z + bx.w1 * v11.z);
    o.w = w0y * (w0x * v00.w + bx.w1 * v01.w) + by.w1 * (w0x * v10.w + bx.w1 * v11.w);
    *reinterpret_cast<float4*>(y + (long long)b * obs + ((long long)oy * Wo + ox) * old_ + c) = o;
  }
}

// A pointwise layer (1x1 conv + activation) followed by a bilinear down-sampling only needs the layer at the four neighbours of every output
// sample: gather those (out[2i + a][2j + b] = x[i_a(i)][j_b(j)], a / b = the lower / upper neighbour), run the layer on the 2Ho x 2Wo gather,
// then blend with the SAME weights and arithmetic order as resize_ac4_kernel - bit-identical to layer-at-full-resolution + resize when the
// layer treats pixels independently.  to_context at the 256^2 scale (appmotioncodebook_arch.py:416-418): a quarter of the pixels.
__global__ void gather_bil4_kernel(const float* __restrict__ x, int Hi, int Wi, int C4, long long ibs, int ild, float* __restrict__ y, int Ho, int Wo, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C4) * 4; long long pp = i / C4; int gx = (int)(pp % (2 * Wo)); long long t = pp / (2 * Wo); int gy = (int)(t % (2 * Ho)); int b = (int)(t / (2 * Ho));
    const Bil by = bil_ac(gy >> 1, Hi, Ho), bx = bil_ac(gx >> 1, Wi, Wo);
    const int iy = (gy & 1) ? by.i1 : by.i0, ix = (gx & 1) ? bx.i1 : bx.i0;
    *reinterpret_cast<float4*>(y + (pp * C4) * 4 + c) = __ldg(reinterpret_cast<const float4*>(x + (long long)b * ibs + ((long long)iy * Wi + ix) * ild + c));
  }
}
__global__ void blend_bil4_kernel(const float* __restrict__ g, int Hi, int Wi, int C4, float* __restrict__ y, int Ho, int Wo, long long obs, int old_, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C4) * 4; long long pp = i / C4; int ox = (int)(pp % Wo); long long t = pp / Wo; int oy = (int)(t % Ho); int b = (int)(t / Ho);
    const Bil by = bil_ac(oy, Hi, Ho), bx = bil_ac(ox, Wi, Wo);
    const float* gb = g + ((((long long)b * 2 * Ho + 2 * oy) * (2 * Wo) + 2 * ox) * C4) * 4 + c;
    const long long rowp = (long long)2 * Wo * C4 * 4;
    const float4 v00 = __ldg(reinterpret_cast<const float4*>(gb)), v01 = __ldg(reinterpret_cast<const float4*>(gb + C4 * 4));
    const float4 v10 = __ldg(reinterpret_cast<const float4*>(gb + rowp)), v11 = __ldg(reinterpret_cast<const float4*>(gb + rowp + C4 * 4));
    const float w0y = 1.f - by.w1, w0x = 1.f - bx.w1;
    float4 o;
    o.x = w0y * (w0x * v00.x + bx.w1 * v01.x) + by.w1 * (w0x * v10.x + bx.w1 * v11.x);
    o.y = w0y * (w0x * v00.y + bx.w1 * v01.y) + by.w1 * (w0x * v10.y + bx.w1 * v11.y);
    o.z = w0y * (w0x * v00.z + bx.w1 * v01.z) + by.w1 * (w0x * v10.z + bx.w1 * v11.z);
    o.w = w0y * (w0x * v00.w + bx.w1 * v01.w) + by.w1 * (w0x * v10.w + bx.w1 * v11.w);
    *reinterpret_cast<float4*>(y + (long long)b * obs + ((long long)oy * Wo + ox) * old_ + c) = o;
  }
}

}  // namespace

extern "C" int sma_gather_bilinear4(const float* x, int B, int Hi, int Wi, int C, int64_t ibs, int ild, float* g, int Ho, int Wo, sma_stream_t stream) {
  if (!x || !g || B <= 0 || Hi <= 0 || Wi <= 0 || C <= 0 || Ho <= 0 || Wo <= 0) return SMA_ERR_BAD_ARG;
  if ((C & 3) || (ild & 3) || (ibs & 3) || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g)) & 15)) return SMA_ERR_UNSUPPORTED;
  long long total = (long long)B * 4 * Ho * Wo * (C / 4);
  long long blocks = (total + 255) / 256; if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  gather_bil4_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, Hi, Wi, C / 4, ibs, ild, g, Ho, Wo, total);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}
extern "C" int sma_blend_bilinear4(const float* g, int B, int Hi, int Wi, int C, float* y, int Ho, int Wo, int64_t obs, int old_, sma_stream_t stream) {
  if (!g || !y || B <= 0 || Hi <= 0 || Wi <= 0 || C <= 0 || Ho <= 0 || Wo <= 0) return SMA_ERR_BAD_ARG;
  if ((C & 3) || (old_ & 3) || (obs & 3) || ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(y)) & 15)) return SMA_ERR_UNSUPPORTED;
  long long total = (long long)B * Ho * Wo * (C / 4);
  long long blocks = (total + 255) / 256; if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  blend_bil4_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(g, Hi, Wi, C / 4, y, Ho, Wo, obs, old_, total);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}

extern "C" int sma_warp_occlude_gather_fwd(const float* feat, int64_t fbs, int B, int H, int W, int C, const float* flow, const float* occ,
                                           int hf, int wf, int Hg, int Wg, float* out, sma_stream_t stream);
extern "C" int sma_warp_occlude_fwd(const float* feat, int64_t fbs, int B, int H, int W, int C, const float* flow, const float* occ,
                                    int hf, int wf, float* out, sma_stream_t stream) {
  if (!feat || !flow || !out || B <= 0 || H <= 1 || W <= 1 || C <= 0 || hf <= 1 || wf <= 1) return SMA_ERR_BAD_ARG;
  if ((C & 3) || (fbs & 3) || ((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(out)) & 15) ||
      (reinterpret_cast<uintptr_t>(flow) & 7))
    return SMA_ERR_UNSUPPORTED;
  return sma_warp_occlude_gather_fwd(feat, fbs, B, H, W, C, flow, occ, hf, wf, 0, 0, out, stream);
}

extern "C" int sma_warp_occlude_gather_fwd(const float* feat, int64_t fbs, int B, int H, int W, int C, const float* flow, const float* occ,
                                           int hf, int wf, int Hg, int Wg, float* out, sma_stream_t stream) {
  if (!feat || !flow || !out || B <= 0 || H <= 1 || W <= 1 || C <= 0 || hf <= 1 || wf <= 1 || Hg < 0 || Wg < 0 || (Hg == 0) != (Wg == 0)) return SMA_ERR_BAD_ARG;
  if ((C & 3) || (fbs & 3) || ((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(out)) & 15) ||
      (reinterpret_cast<uintptr_t>(flow) & 7))
    return SMA_ERR_UNSUPPORTED;
  const bool v8 = (C & 7) == 0 && (fbs & 7) == 0 && ((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(out)) & 31) == 0;
  long long total4 = (long long)B * (Hg ? 4LL * Hg * Wg : (long long)H * W) * (C / (v8 ? 8 : 4));
  long long blocks = (total4 + 255) / 256; if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  if (v8) warp_occlude_kernel<8><<<(int)blocks, 256, 0, as_stream(stream)>>>(feat, fbs, B, H, W, C, flow, occ, hf, wf, out, total4, Hg, Wg);
  else warp_occlude_kernel<4><<<(int)blocks, 256, 0, as_stream(stream)>>>(feat, fbs, B, H, W, C, flow, occ, hf, wf, out, total4, Hg, Wg);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}

extern "C" int sma_resize_bilinear_ac(const float* x, int B, int Hi, int Wi, int C, int64_t ibs, int ild, float* y, int Ho, int Wo,
                                      int64_t obs, int old_, sma_stream_t stream) {
  if (!x || !y || B <= 0 || Hi <= 0 || Wi <= 0 || C <= 0 || Ho <= 0 || Wo <= 0) return SMA_ERR_BAD_ARG;
  const bool v4 = (C & 3) == 0 && (ild & 3) == 0 && (old_ & 3) == 0 && (ibs & 3) == 0 && (obs & 3) == 0 &&
                  ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  long long total = (long long)B * Ho * Wo * (v4 ? C / 4 : C);
  long long blocks = (total + 255) / 256; if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  if (v4) resize_ac4_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, Hi, Wi, C / 4, ibs, ild, y, Ho, Wo, obs, old_, total);
  else resize_ac_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, Hi, Wi, C, ibs, ild, y, Ho, Wo, obs, old_, total);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}
