// Motion-estimator heads and the small glue kernels of the per-frame path (all HBM/latency-bound,
// CUDA-core work; none of it is a dense contraction).
#include "sma_common.cuh"
#include <math_constants.h>

namespace {

// 2*(i/(n-1))-1 exactly as make_coordinate_grid evaluates it (utils/motion_estimator_util.py:63-64)
__device__ __forceinline__ float grid_coord(int i, int n) { return 2.f * ((float)i / (float)(n - 1)) - 1.f; }

__global__ void antialias_down4_kernel(const float* __restrict__ x, int B, int C, int H, int W, const float* __restrict__ k13,
                                       float* __restrict__ y, int yld) {
  __shared__ float ks[169];
  for (int i = threadIdx.x; i < 169; i += blockDim.x) ks[i] = k13[i];
  __syncthreads();
  const int Ho = H / 4, Wo = W / 4;
  long long total = (long long)B * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long pp = i / C; int ox = (int)(pp % Wo); long long t = pp / Wo; int oy = (int)(t % Ho); int b = (int)(t / Ho);
    const float* xc = x + ((long long)b * C + c) * H * W;
    float acc = 0.f;
    for (int ky = 0; ky < 13; ky++) {
      int iy = oy * 4 + ky - 6; if ((unsigned)iy >= (unsigned)H) continue;
      for (int kx = 0; kx < 13; kx++) {
        int ix = ox * 4 + kx - 6; if ((unsigned)ix >= (unsigned)W) continue;
        acc = fmaf(ks[ky * 13 + kx], __ldg(xc + (long long)iy * W + ix), acc);
      }
    }
    y[pp * yld + c] = acc;
  }
}

__global__ void avgpool2_kernel(const float* __restrict__ x, int B, int H, int W, int C, float* __restrict__ y, int yld) {
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  long long total = (long long)B * Ho * Wo * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4); long long pp = i / C4; int ox = (int)(pp % Wo); long long t = pp / Wo; int oy = (int)(t % Ho); int b = (int)(t / Ho);
    const float* p = x + (((long long)b * H + oy * 2) * W + ox * 2) * C + c4 * 4;
    float4 a = __ldg(reinterpret_cast<const float4*>(p)), bq = __ldg(reinterpret_cast<const float4*>(p + C));
    float4 c = __ldg(reinterpret_cast<const float4*>(p + (long long)W * C)), d = __ldg(reinterpret_cast<const float4*>(p + (long long)W * C + C));
    float4 o = make_float4((a.x + bq.x + c.x + d.x) * 0.25f, (a.y + bq.y + c.y + d.y) * 0.25f, (a.z + bq.z + c.z + d.z) * 0.25f, (a.w + bq.w + c.w + d.w) * 0.25f);
    *reinterpret_cast<float4*>(y + pp * yld + c4 * 4) = o;
  }
}

__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = is_max ? -CUDART_INF_F : 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];
  return r;
}

// one block per (k, b): softmax(logit / T) over h*w, expectation of the grid and of the 4 jacobian maps
__global__ void kp_head_kernel(const float* __restrict__ pred, int h, int w, int ld, int K, float T, float* __restrict__ value, float* __restrict__ jac) {
  __shared__ float sh[32];
  const int k = blockIdx.x, b = blockIdx.y, n = h * w;
  const float* pb = pred + (long long)b * n * ld;
  float mx = -CUDART_INF_F;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, __ldg(pb + (long long)i * ld + k) / T);
  mx = block_reduce(mx, sh, true);
  float s = 0.f, sx = 0.f, sy = 0.f, j0 = 0.f, j1 = 0.f, j2 = 0.f, j3 = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float* px = pb + (long long)i * ld;
    float e = expf(__ldg(px + k) / T - mx);
    int yy = i / w, xx = i - yy * w;
    s += e; sx = fmaf(e, grid_coord(xx, w), sx); sy = fmaf(e, grid_coord(yy, h), sy);
    if (jac) {
      const float* pj = px + K + k * 4;
      j0 = fmaf(e, __ldg(pj), j0); j1 = fmaf(e, __ldg(pj + 1), j1); j2 = fmaf(e, __ldg(pj + 2), j2); j3 = fmaf(e, __ldg(pj + 3), j3);
    }
  }
  s = block_reduce(s, sh, false); sx = block_reduce(sx, sh, false); sy = block_reduce(sy, sh, false);
  if (jac) { j0 = block_reduce(j0, sh, false); j1 = block_reduce(j1, sh, false); j2 = block_reduce(j2, sh, false); j3 = block_reduce(j3, sh, false); }
  if (threadIdx.x == 0) {
    float inv = 1.f / s;
    float* v = value + ((long long)b * K + k) * 2; v[0] = sx * inv; v[1] = sy * inv;
    if (jac) { float* j = jac + ((long long)b * K + k) * 4; j[0] = j0 * inv; j[1] = j1 * inv; j[2] = j2 * inv; j[3] = j3 * inv; }
  }
}

__device__ __forceinline__ void inv2(const float* a, float* o) {
  float det = a[0] * a[3] - a[1] * a[2]; float r = 1.f / det;
  o[0] = a[3] * r; o[1] = -a[1] * r; o[2] = -a[2] * r; o[3] = a[0] * r;
}
__device__ __forceinline__ void mm2(const float* a, const float* b, float* o) {
  o[0] = a[0] * b[0] + a[1] * b[2]; o[1] = a[0] * b[1] + a[1] * b[3];
  o[2] = a[2] * b[0] + a[3] * b[2]; o[3] = a[2] * b[1] + a[3] * b[3];
}

// area of the convex hull of n <= 64 points (monotone chain + shoelace, fp64): what scipy's ConvexHull(points).volume returns for 2-D input
__device__ double hull_area_dev(const float* pts, int n) {
  double x[64], y[64]; int h[130];
  for (int i = 0; i < n; i++) { x[i] = pts[2 * i]; y[i] = pts[2 * i + 1]; }
  for (int i = 1; i < n; i++) {            // insertion sort, lexicographic
    double xi = x[i], yi = y[i]; int j = i - 1;
    while (j >= 0 && (x[j] > xi || (x[j] == xi && y[j] > yi))) { x[j + 1] = x[j]; y[j + 1] = y[j]; j--; }
    x[j + 1] = xi; y[j + 1] = yi;
  }
  if (n < 3) return 0.0;
  int m = 0;
  for (int i = 0; i < n; i++) {            // lower hull
    while (m >= 2 && (x[h[m - 1]] - x[h[m - 2]]) * (y[i] - y[h[m - 2]]) - (y[h[m - 1]] - y[h[m - 2]]) * (x[i] - x[h[m - 2]]) <= 0.0) m--;
    h[m++] = i;
  }
  for (int i = n - 2, t = m + 1; i >= 0; i--) {   // upper hull
    while (m >= t && (x[h[m - 1]] - x[h[m - 2]]) * (y[i] - y[h[m - 2]]) - (y[h[m - 1]] - y[h[m - 2]]) * (x[i] - x[h[m - 2]]) <= 0.0) m--;
    h[m++] = i;
  }
  m--;                                     // the last point repeats the first
  double a = 0.0;
  for (int i = 0; i < m; i++) { int j = (i + 1) % m; a += x[h[i]] * y[h[j]] - x[h[j]] * y[h[i]]; }
  return 0.5 * fabs(a);
}
// movement scale of normalize_kp: sqrt(hull area of the source key-points) / sqrt(hull area of the initial driving key-points), demo.py:26-29
__global__ void hull_scale_kernel(const float* src_v, const float* drv0_v, int K, float* scale_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) scale_out[0] = (float)(sqrt(hull_area_dev(src_v, K)) / sqrt(hull_area_dev(drv0_v, K)));
}

__global__ void normalize_kp_kernel(const float* src_v, const float* src_j, const float* drv_v, const float* drv_j, const float* drv0_v,
                                    const float* drv0_j, int B, int K, float scale, const float* scale_dev, int relative, float* out_v, float* out_j) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K) return;
  int k = i % K;
  if (scale_dev) scale = scale_dev[0];
  if (!relative) {
    out_v[i * 2] = drv_v[i * 2]; out_v[i * 2 + 1] = drv_v[i * 2 + 1];
    for (int j = 0; j < 4; j++) out_j[i * 4 + j] = drv_j[i * 4 + j];
    return;
  }
  out_v[i * 2] = (drv_v[i * 2] - drv0_v[k * 2]) * scale + src_v[k * 2];
  out_v[i * 2 + 1] = (drv_v[i * 2 + 1] - drv0_v[k * 2 + 1]) * scale + src_v[k * 2 + 1];
  float inv0[4], t[4], o[4];
  inv2(drv0_j + k * 4, inv0); mm2(drv_j + i * 4, inv0, t); mm2(t, src_j + k * 4, o);
  for (int j = 0; j < 4; j++) out_j[i * 4 + j] = o[j];
}

// sparse motion k (1..K) at grid point (gx,gy):  J_s J_d^-1 (g - kp_d) + kp_s   (dense_motion_arch.py:84-104)
struct KpXform { float a[4]; float dx, dy, sx, sy; };
__device__ __forceinline__ KpXform make_xform(const float* sv, const float* sj, const float* dv, const float* dj) {
  KpXform t; float inv[4]; inv2(dj, inv); mm2(sj, inv, t.a); t.dx = dv[0]; t.dy = dv[1]; t.sx = sv[0]; t.sy = sv[1]; return t;
}
__device__ __forceinline__ void apply_xform(const KpXform& t, float gx, float gy, float& ox, float& oy) {
  float zx = gx - t.dx, zy = gy - t.dy;
  ox = t.a[0] * zx + t.a[1] * zy + t.sx; oy = t.a[2] * zx + t.a[3] * zy + t.sy;
}

// thread per (b, pixel, k)
__global__ void dense_motion_prep_kernel(const float* __restrict__ src, int h, int w, const float* __restrict__ sv, const float* __restrict__ sj,
                                         const float* __restrict__ dv, const float* __restrict__ dj, int B, int K, float var,
                                         float* __restrict__ hg, int hg_ld, float* __restrict__ heat_out) {
  const int K1 = K + 1;
  long long total = (long long)B * h * w * K1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i % K1); long long pp = i / K1; int x = (int)(pp % w); long long t = pp / w; int y = (int)(t % h); int b = (int)(t / h);
    float gx = grid_coord(x, w), gy = grid_coord(y, h);
    float heat = 0.f, mx = gx, my = gy;
    if (k > 0) {
      const float* dvk = dv + ((long long)b * K + k - 1) * 2; const float* svk = sv + (k - 1) * 2;
      float ddx = gx - dvk[0], ddy = gy - dvk[1], sdx = gx - svk[0], sdy = gy - svk[1];
      float gd = expf(-0.5f * (ddx * ddx + ddy * ddy) / var), gs = expf(-0.5f * (sdx * sdx + sdy * sdy) / var);
      heat = gd - gs;
      if (heat_out) heat_out[pp * K + (k - 1)] = gd;
      KpXform xf = make_xform(svk, sj + (k - 1) * 4, dvk, dj + ((long long)b * K + k - 1) * 4);
      apply_xform(xf, gx, gy, mx, my);
    }
    // grid_sample bilinear / zeros / align_corners=False on the (h,w,3) source
    float ix = ((mx + 1.f) * (float)w - 1.f) * 0.5f, iy = ((my + 1.f) * (float)h - 1.f) * 0.5f;
    float fx = floorf(ix), fy = floorf(iy);
    int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    float r = 0.f, g = 0.f, bl = 0.f;
    bool vx0 = (unsigned)x0 < (unsigned)w, vx1 = (unsigned)x1 < (unsigned)w, vy0 = (unsigned)y0 < (unsigned)h, vy1 = (unsigned)y1 < (unsigned)h;
    if (vy0 && vx0) { const float* p = src + ((long long)y0 * w + x0) * 3; float ww = wy0 * wx0; r = fmaf(ww, p[0], r); g = fmaf(ww, p[1], g); bl = fmaf(ww, p[2], bl); }
    if (vy0 && vx1) { const float* p = src + ((long long)y0 * w + x1) * 3; float ww = wy0 * wx1; r = fmaf(ww, p[0], r); g = fmaf(ww, p[1], g); bl = fmaf(ww, p[2], bl); }
    if (vy1 && vx0) { const float* p = src + ((long long)y1 * w + x0) * 3; float ww = wy1 * wx0; r = fmaf(ww, p[0], r); g = fmaf(ww, p[1], g); bl = fmaf(ww, p[2], bl); }
    if (vy1 && vx1) { const float* p = src + ((long long)y1 * w + x1) * 3; float ww = wy1 * wx1; r = fmaf(ww, p[0], r); g = fmaf(ww, p[1], g); bl = fmaf(ww, p[2], bl); }
    *reinterpret_cast<float4*>(hg + pp * hg_ld + k * 4) = make_float4(heat, r, g, bl);
  }
}

// thread per (b, pixel); K+1 <= 32
__global__ void dense_motion_head_kernel(const float* __restrict__ logits, int ld, int h, int w, const float* __restrict__ sv, const float* __restrict__ sj,
                                         const float* __restrict__ dv, const float* __restrict__ dj, int B, int K, float* __restrict__ deform,
                                         float* __restrict__ occ, float* __restrict__ mask_out) {
  const int K1 = K + 1;
  long long total = (long long)B * h * w;
  for (long long pp = blockIdx.x * (long long)blockDim.x + threadIdx.x; pp < total; pp += (long long)gridDim.x * blockDim.x) {
    int x = (int)(pp % w); long long t = pp / w; int y = (int)(t % h); int b = (int)(t / h);
    const float* lg = logits + pp * ld;
    float mx = -CUDART_INF_F;
    for (int k = 0; k < K1; k++) mx = fmaxf(mx, __ldg(lg + k));
    float s = 0.f;
    for (int k = 0; k < K1; k++) s += expf(__ldg(lg + k) - mx);
    float inv = 1.f / s;
    float gx = grid_coord(x, w), gy = grid_coord(y, h);
    float ax = 0.f, ay = 0.f;
    for (int k = 0; k < K1; k++) {
      float m = expf(__ldg(lg + k) - mx) * inv;
      if (mask_out) mask_out[pp * K1 + k] = m;
      float px = gx, py = gy;
      if (k > 0) {
        KpXform xf = make_xform(sv + (k - 1) * 2, sj + (k - 1) * 4, dv + ((long long)b * K + k - 1) * 2, dj + ((long long)b * K + k - 1) * 4);
        apply_xform(xf, gx, gy, px, py);
      }
      ax = fmaf(m, px, ax); ay = fmaf(m, py, ay);
    }
    deform[pp * 2] = ax; deform[pp * 2 + 1] = ay;
    if (occ) occ[pp] = 1.f / (1.f + expf(-__ldg(lg + K1)));
  }
}

// torch.linspace(-1, 1, n)[i] as ATen evaluates it (symmetric two-sided formula)
__device__ __forceinline__ float linspace_pm1(int i, int n) {
  float step = 2.f / (float)(n - 1);
  return i < n / 2 ? -1.f + step * (float)i : 1.f - step * (float)(n - 1 - i);
}

__global__ void flow_to_px_kernel(const float* __restrict__ m, int B, int h, int w, float* __restrict__ o, int ld) {
  long long total = (long long)B * h * w;
  const float hx = ((float)h - 1.f) * 0.5f;   // the reference scales both components by (shape[1]-1)/2
  for (long long pp = blockIdx.x * (long long)blockDim.x + threadIdx.x; pp < total; pp += (long long)gridDim.x * blockDim.x) {
    int x = (int)(pp % w); int y = (int)((pp / w) % h);
    o[pp * ld] = (m[pp * 2] - linspace_pm1(x, h)) * hx;
    o[pp * ld + 1] = (m[pp * 2 + 1] - linspace_pm1(y, w)) * hx;
  }
}

__global__ void flow_update_kernel(const float* __restrict__ m, const float* __restrict__ occ, const float* __restrict__ res, int ld, int B, int h,
                                   int w, float* __restrict__ mo, float* __restrict__ oo) {
  long long total = (long long)B * h * w;
  const float hx = ((float)h - 1.f) * 0.5f;
  for (long long pp = blockIdx.x * (long long)blockDim.x + threadIdx.x; pp < total; pp += (long long)gridDim.x * blockDim.x) {
    const float* r = res + pp * ld;
    mo[pp * 2] = m[pp * 2] + r[0] / hx; mo[pp * 2 + 1] = m[pp * 2 + 1] + r[1] / hx;
    oo[pp] = 1.f / (1.f + expf(-(occ[pp] + r[2])));
  }
}

__global__ void ignore_mask_kernel(const float* __restrict__ m, int B, int h, int w, int ht, int wt, uint8_t* __restrict__ mask) {
  long long total = (long long)B * ht * wt;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int ox = (int)(i % wt); int oy = (int)((i / wt) % ht); int b = (int)(i / ((long long)wt * ht));
    float sy = (ht > 1 ? (float)(h - 1) / (float)(ht - 1) : 0.f) * oy, sx = (wt > 1 ? (float)(w - 1) / (float)(wt - 1) : 0.f) * ox;
    int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1); int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
    float ly = sy - y0, lx = sx - x0;
    const float* mb = m + (long long)b * h * w * 2;
    bool ig = false;
    for (int c = 0; c < 2; c++) {
      float v = (1.f - ly) * ((1.f - lx) * mb[(y0 * w + x0) * 2 + c] + lx * mb[(y0 * w + x1) * 2 + c]) +
                ly * ((1.f - lx) * mb[(y1 * w + x0) * 2 + c] + lx * mb[(y1 * w + x1) * 2 + c]);
      ig = ig || v > 1.f || v < -1.f;
    }
    mask[i] = ig ? 1 : 0;
  }
}

__global__ void sft_combine_kernel(const float4* __restrict__ dec, const float4* __restrict__ sc, const float4* __restrict__ sh, float w, long long n4,
                                   float4* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 d = __ldg(dec + i), s = __ldg(sc + i), h = __ldg(sh + i), o;
    o.x = d.x + w * (d.x * s.x + h.x); o.y = d.y + w * (d.y * s.y + h.y); o.z = d.z + w * (d.z * s.z + h.z); o.w = d.w + w * (d.w * s.w + h.w);
    out[i] = o;
  }
}

__global__ void to_uint8_kernel(const float* __restrict__ x, long long npix, int C, int ld, int bgr, uint8_t* __restrict__ out) {
  long long total = npix * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long pp = i / C;
    float v = x[pp * ld + (bgr ? C - 1 - c : c)];
    v = fminf(fmaxf(v, -1.f), 1.f);
    v = (v + 1.f) / 2.f;
    out[i] = (uint8_t)rintf(v * 255.f);
  }
}

// uint8 HWC frames -> normalised fp32 NCHW, bit-exact with the reference's host-side preparation:
// img.astype(float32) / 255. (demo.py:180-181), HWC -> CHW with optional BGR -> RGB (img_util.py:13-39), normalize(mean 0.5, std 0.5) (demo.py:183-185)
__global__ void u8hwc_to_f32nchw_kernel(const uint8_t* __restrict__ x, int B, int C, int HW, int swap_rb, float* __restrict__ y) {
  long long total = (long long)B * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int p = (int)(i % HW); long long t = i / HW; int c = (int)(t % C); int b = (int)(t / C);
    float v = (float)x[((long long)b * HW + p) * C + (swap_rb ? C - 1 - c : c)] / 255.f;
    y[i] = (v - 0.5f) / 0.5f;
  }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int B, int C, int HW, float* __restrict__ y, int yld) {
  long long total = (long long)B * HW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long pp = i / C; int p = (int)(pp % HW); int b = (int)(pp / HW);
    y[pp * yld + c] = x[((long long)b * C + c) * HW + p];
  }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, int B, int C, int HW, int xld, float* __restrict__ y) {
  long long total = (long long)B * HW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int p = (int)(i % HW); long long t = i / HW; int c = (int)(t % C); int b = (int)(t / C);
    y[i] = x[((long long)b * HW + p) * xld + c];
  }
}

inline int nblocks(long long total, int per = 256) { long long b = (total + per - 1) / per; if (b > kNumSMs * 32) b = kNumSMs * 32; if (b < 1) b = 1; return (int)b; }

}  // namespace

extern "C" int sma_antialias_down4(const float* x, int B, int C, int H, int W, const float* k13, float* y, int yld, sma_stream_t s) {
  if (!x || !k13 || !y || B <= 0 || C <= 0 || H < 4 || W < 4 || (H & 3) || (W & 3) || yld < C) return SMA_ERR_BAD_ARG;
  antialias_down4_kernel<<<nblocks((long long)B * (H / 4) * (W / 4) * C), 256, 0, as_stream(s)>>>(x, B, C, H, W, k13, y, yld);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_avgpool2(const float* x, int B, int H, int W, int C, float* y, int yld, sma_stream_t s) {
  if (!x || !y || B <= 0 || H < 2 || W < 2 || (H & 1) || (W & 1)) return SMA_ERR_BAD_ARG;
  if ((C & 3) || (yld & 3) || yld < C) return SMA_ERR_UNSUPPORTED;
  avgpool2_kernel<<<nblocks((long long)B * (H / 2) * (W / 2) * (C / 4)), 256, 0, as_stream(s)>>>(x, B, H, W, C, y, yld);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_kp_head_fwd(const float* pred, int B, int h, int w, int ld, int K, float T, float* value, float* jac, sma_stream_t s) {
  if (!pred || !value || B <= 0 || h <= 1 || w <= 1 || K <= 0 || ld < (jac ? 5 * K : K) || T <= 0.f) return SMA_ERR_BAD_ARG;
  kp_head_kernel<<<dim3(K, B), 256, 0, as_stream(s)>>>(pred, h, w, ld, K, T, value, jac);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_normalize_kp(const float* sv, const float* sj, const float* dv, const float* dj, const float* d0v, const float* d0j, int B,
                                int K, float scale, const float* scale_dev, int relative, float* ov, float* oj, sma_stream_t s) {
  if (!sv || !sj || !dv || !dj || !d0v || !d0j || !ov || !oj || B <= 0 || K <= 0) return SMA_ERR_BAD_ARG;
  normalize_kp_kernel<<<cdiv(B * K, 128), 128, 0, as_stream(s)>>>(sv, sj, dv, dj, d0v, d0j, B, K, scale, scale_dev, relative, ov, oj);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_hull_scale(const float* src_v, const float* drv0_v, int K, float* scale_out, sma_stream_t s) {
  if (!src_v || !drv0_v || !scale_out || K < 3 || K > 64) return SMA_ERR_BAD_ARG;
  hull_scale_kernel<<<1, 32, 0, as_stream(s)>>>(src_v, drv0_v, K, scale_out);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_u8hwc_to_f32nchw(const uint8_t* x, int B, int H, int W, int C, int swap_rb, float* y, sma_stream_t s) {
  if (!x || !y || B <= 0 || H <= 0 || W <= 0 || C <= 0) return SMA_ERR_BAD_ARG;
  u8hwc_to_f32nchw_kernel<<<nblocks((long long)B * C * H * W), 256, 0, as_stream(s)>>>(x, B, C, H * W, swap_rb, y);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_dense_motion_prep(const float* src64, int h, int w, const float* sv, const float* sj, const float* dv, const float* dj, int B,
                                     int K, float var, float* hg, int hg_ld, float* heat, sma_stream_t s) {
  if (!src64 || !sv || !sj || !dv || !dj || !hg || B <= 0 || K <= 0 || h <= 1 || w <= 1 || var <= 0.f) return SMA_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(hg) & 15) || (hg_ld & 3) || hg_ld < 4 * (K + 1)) return SMA_ERR_UNSUPPORTED;
  dense_motion_prep_kernel<<<nblocks((long long)B * h * w * (K + 1)), 256, 0, as_stream(s)>>>(src64, h, w, sv, sj, dv, dj, B, K, var, hg, hg_ld, heat);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_dense_motion_head(const float* logits, int ld, int h, int w, const float* sv, const float* sj, const float* dv, const float* dj,
                                     int B, int K, float* deform, float* occ, float* mask_out, sma_stream_t s) {
  if (!logits || !sv || !sj || !dv || !dj || !deform || B <= 0 || K <= 0 || ld < K + (occ ? 2 : 1)) return SMA_ERR_BAD_ARG;
  dense_motion_head_kernel<<<nblocks((long long)B * h * w, 128), 128, 0, as_stream(s)>>>(logits, ld, h, w, sv, sj, dv, dj, B, K, deform, occ, mask_out);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
// im2col of a few-channel map (the 2-channel pixel-unit flow feeding the 7x7 BasicMotionEncoder.convf1, appmotioncodebook_arch.py:136,142): as an
// implicit GEMM the conv would contract over kh*kw taps of a 32-channel zero-padded chunk each (1568 K values for 98 real ones); unfolded to
// K = k*k*C (padded to `Kp`) it is ONE 128-deep 1x1 conv on the tensor cores.  out[b][y][x][(ky*k+kx)*C + c] = x[b][y+ky-pad][x+kx-pad][c], zero
// outside the image and for columns >= k*k*C.  One thread per (pixel, 4 columns).
// block = 8 x 32 output pixels of one frame: the (8 + k - 1) x (32 + k - 1) x C input halo is staged in shared memory once, then every thread writes
// float4 groups of the unfolded rows (the kernel is bound by the HBM writes of the (B,H,W,Kp) result)
__global__ void __launch_bounds__(256) im2col_small_kernel(const float* __restrict__ x, int B, int H, int W, int ld, int C, int k, int pad, float* __restrict__ out, int Kp,
                                                         int tiles_x, int tiles_y) {
  extern __shared__ float sh[];                       // [(8 + k - 1)][(32 + k - 1)][C]
  const int t = blockIdx.x; const int b = t / (tiles_x * tiles_y); const int tr = t - b * tiles_x * tiles_y;
  const int ty0 = (tr / tiles_x) * 8, tx0 = (tr % tiles_x) * 32;
  const int hw = 32 + k - 1, hh = 8 + k - 1;
  for (int i = threadIdx.x; i < hh * hw * C; i += 256) {
    const int c = i % C; const int px = (i / C) % hw; const int py = i / (C * hw);
    const int iy = ty0 + py - pad, ix = tx0 + px - pad;
    sh[i] = ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W) ? __ldg(x + (((long long)b * H + iy) * W + ix) * ld + c) : 0.f;
  }
  // column -> offset inside the staged halo (relative to the pixel's top-left tap), -1 for the zero padding columns: one table per block
  __shared__ __align__(16) int off[256];
  const int KK = k * k * C, kq = Kp >> 2;
  for (int col = threadIdx.x; col < Kp && col < 256; col += 256) {
    int o = -1;
    if (col < KK) { const int tap = col / C, c = col - tap * C; const int ky = tap / k, kx = tap - ky * k; o = (ky * hw + kx) * C + c; }
    off[col] = o;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256 * kq; i += 256) {
    const int q = i % kq; const int p = i / kq; const int py = p >> 5, px = p & 31;
    const int oy = ty0 + py, ox = tx0 + px;
    if (oy >= H || ox >= W) continue;
    const float* base = sh + (py * hw + px) * C;
    const int4 o4 = *reinterpret_cast<const int4*>(&off[q * 4]);
    const float4 v = make_float4(o4.x >= 0 ? base[o4.x] : 0.f, o4.y >= 0 ? base[o4.y] : 0.f, o4.z >= 0 ? base[o4.z] : 0.f, o4.w >= 0 ? base[o4.w] : 0.f);
    *reinterpret_cast<float4*>(out + (((long long)b * H + oy) * W + ox) * Kp + q * 4) = v;
  }
}
// A k x k convolution with very few output channels (the 3-channel image / [delta-flow | delta-occlusion] heads) as a pointwise layer + this sum:
// the tensor-core kernel pays the same fixed cost per MMA whether it produces 8 or 256 columns, so the k*k taps are made COLUMNS of a 1x1 conv
// (P[pixel][tap * C + c] = sum_ci w[c][ci][tap] x[pixel][ci], k*k times fewer MMAs) and the taps are gathered here:
// out[y][x][c] = bias[c] + sum_tap P[y + ky - pad][x + kx - pad][tap * C + c]; taps outside the map are skipped (= the conv's zero padding).
template <int C>
__global__ void tapsum_kernel(const float* __restrict__ P, int ldp, int B, int H, int W, int k, int pad, const float* __restrict__ bias,
                              float* __restrict__ out, int ldo) {
  const long long n = (long long)B * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long b = i / ((long long)W * H);
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] = bias ? __ldg(bias + c) : 0.f;
    for (int ky = 0; ky < k; ky++) {
      const int yy = y + ky - pad;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < k; kx++) {
        const int xx = x + kx - pad;
        if (xx < 0 || xx >= W) continue;
        const float* src = P + ((b * H + yy) * W + xx) * ldp + (ky * k + kx) * C;
#pragma unroll
        for (int c = 0; c < C; c++) acc[c] += __ldg(src + c);
      }
    }
#pragma unroll
    for (int c = 0; c < C; c++) out[i * ldo + c] = acc[c];
  }
}
// The 3x3 / pad 1 case through shared memory: a block owns an 8 x 32 tile of outputs, reads the 10 x 34 pixels it needs ONCE with 128-bit loads (the
// per-thread form above fetches 12 bytes out of 27 different 128-byte rows and is bound by the L1 request rate: 1.3 TB/s on the 256^2 head), keeps them
// column-major ([tap column][pixel], odd pitch) and sums the taps in the same order (bias, then ky, kx ascending; a tap outside the map adds 0).
constexpr int TS_TH = 8, TS_TW = 32, TS_PW = TS_TW + 2, TS_NP = (TS_TH + 2) * TS_PW, TS_PITCH = 345;
template <int C>
__global__ void __launch_bounds__(256) tapsum3_tiled_kernel(const float* __restrict__ P, int ldp, int B, int H, int W, const float* __restrict__ bias,
                                                            float* __restrict__ out, int ldo, int tiles_x, int tiles_y) {
  constexpr int NQ = (9 * C + 3) / 4;
  __shared__ float sm[4 * NQ * TS_PITCH];
  int t = blockIdx.x;
  const int tx0 = (t % tiles_x) * TS_TW; t /= tiles_x;
  const int ty0 = (t % tiles_y) * TS_TH; const long long b = t / tiles_y;
  for (int i = threadIdx.x; i < TS_NP * NQ; i += 256) {
    const int p = i / NQ, q = i - p * NQ;
    const int py = p / TS_PW, px = p - py * TS_PW;
    const int yy = ty0 + py - 1, xx = tx0 + px - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(reinterpret_cast<const float4*>(P + ((b * H + yy) * W + xx) * ldp) + q);
    float* d = sm + (4 * q) * TS_PITCH + p;
    d[0] = v.x; d[TS_PITCH] = v.y; d[2 * TS_PITCH] = v.z; d[3 * TS_PITCH] = v.w;
  }
  __syncthreads();
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int x = tx0 + lx, y = ty0 + ly;
  if (x >= W || y >= H) return;
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; c++) acc[c] = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ky++)
#pragma unroll
    for (int kx = 0; kx < 3; kx++) {
      const float* src = sm + ((ky * 3 + kx) * C) * TS_PITCH + (ly + ky) * TS_PW + lx + kx;
#pragma unroll
      for (int c = 0; c < C; c++) acc[c] += src[c * TS_PITCH];
    }
  float* o = out + ((b * H + y) * W + x) * ldo;
#pragma unroll
  for (int c = 0; c < C; c++) o[c] = acc[c];
}
extern "C" int sma_im2col_small(const float* x, int B, int H, int W, int ld, int C, int k, int pad, float* out, int Kp, sma_stream_t s) {
  if (!x || !out || B <= 0 || H <= 0 || W <= 0 || C <= 0 || k <= 0 || pad < 0 || ld < C || (Kp & 3) || Kp < k * k * C) return SMA_ERR_BAD_ARG;
  if (reinterpret_cast<uintptr_t>(out) & 15) return SMA_ERR_BAD_ARG;
  const int smem = (8 + k - 1) * (32 + k - 1) * C * (int)sizeof(float);
  if (smem > 40 * 1024 || Kp > 256) return SMA_ERR_UNSUPPORTED;
  const int tiles_x = (W + 31) / 32, tiles_y = (H + 7) / 8;
  im2col_small_kernel<<<B * tiles_x * tiles_y, 256, smem, as_stream(s)>>>(x, B, H, W, ld, C, k, pad, out, Kp, tiles_x, tiles_y);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_conv_tapsum(const float* P, int ldp, int B, int H, int W, int C, int k, int pad, const float* bias, float* out, int ldo, sma_stream_t s) {
  if (!P || !out || B <= 0 || H <= 0 || W <= 0 || k <= 0 || pad < 0 || C <= 0 || ldp < k * k * C || ldo < C) return SMA_ERR_BAD_ARG;
  if (C > 4) return SMA_ERR_UNSUPPORTED;
  if (k == 3 && pad == 1 && C <= 3 && (ldp & 3) == 0 && ldp >= 4 * ((9 * C + 3) / 4) && (reinterpret_cast<uintptr_t>(P) & 15) == 0) {
    const int tiles_x = (W + TS_TW - 1) / TS_TW, tiles_y = (H + TS_TH - 1) / TS_TH;
    const long long nb = (long long)B * tiles_x * tiles_y;
    if (nb <= 0x7fffffffLL) {
      const int g = (int)nb;
      if (C == 1) tapsum3_tiled_kernel<1><<<g, 256, 0, as_stream(s)>>>(P, ldp, B, H, W, bias, out, ldo, tiles_x, tiles_y);
      else if (C == 2) tapsum3_tiled_kernel<2><<<g, 256, 0, as_stream(s)>>>(P, ldp, B, H, W, bias, out, ldo, tiles_x, tiles_y);
      else tapsum3_tiled_kernel<3><<<g, 256, 0, as_stream(s)>>>(P, ldp, B, H, W, bias, out, ldo, tiles_x, tiles_y);
      SMA_LAUNCH_CHECK(); return SMA_OK;
    }
  }
  const long long n = (long long)B * H * W;
  switch (C) {
    case 1: tapsum_kernel<1><<<nblocks(n), 256, 0, as_stream(s)>>>(P, ldp, B, H, W, k, pad, bias, out, ldo); break;
    case 2: tapsum_kernel<2><<<nblocks(n), 256, 0, as_stream(s)>>>(P, ldp, B, H, W, k, pad, bias, out, ldo); break;
    case 3: tapsum_kernel<3><<<nblocks(n), 256, 0, as_stream(s)>>>(P, ldp, B, H, W, k, pad, bias, out, ldo); break;
    default: tapsum_kernel<4><<<nblocks(n), 256, 0, as_stream(s)>>>(P, ldp, B, H, W, k, pad, bias, out, ldo); break;
  }
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_flow_to_px(const float* m, int B, int h, int w, float* o, int ld, sma_stream_t s) {
  if (!m || !o || B <= 0 || h <= 1 || w <= 1 || ld < 2) return SMA_ERR_BAD_ARG;
  flow_to_px_kernel<<<nblocks((long long)B * h * w), 256, 0, as_stream(s)>>>(m, B, h, w, o, ld);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_flow_update(const float* m, const float* occ, const float* res, int ld, int B, int h, int w, float* mo, float* oo, sma_stream_t s) {
  if (!m || !occ || !res || !mo || !oo || B <= 0 || h <= 1 || w <= 1 || ld < 3) return SMA_ERR_BAD_ARG;
  flow_update_kernel<<<nblocks((long long)B * h * w), 256, 0, as_stream(s)>>>(m, occ, res, ld, B, h, w, mo, oo);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_motion_ignore_mask(const float* m, int B, int h, int w, int ht, int wt, uint8_t* mask, sma_stream_t s) {
  if (!m || !mask || B <= 0 || h <= 0 || w <= 0 || ht <= 0 || wt <= 0) return SMA_ERR_BAD_ARG;
  ignore_mask_kernel<<<nblocks((long long)B * ht * wt), 256, 0, as_stream(s)>>>(m, B, h, w, ht, wt, mask);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_sft_combine(const float* dec, const float* sc, const float* sh, float w, int64_t n, float* out, sma_stream_t s) {
  if (!dec || !sc || !sh || !out || n <= 0) return SMA_ERR_BAD_ARG;
  if ((n & 3) || ((reinterpret_cast<uintptr_t>(dec) | reinterpret_cast<uintptr_t>(sc) | reinterpret_cast<uintptr_t>(sh) | reinterpret_cast<uintptr_t>(out)) & 15))
    return SMA_ERR_UNSUPPORTED;
  sft_combine_kernel<<<nblocks(n / 4), 256, 0, as_stream(s)>>>(reinterpret_cast<const float4*>(dec), reinterpret_cast<const float4*>(sc),
                                                               reinterpret_cast<const float4*>(sh), w, n / 4, reinterpret_cast<float4*>(out));
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_to_uint8(const float* x, int B, int H, int W, int C, int ld, int bgr, uint8_t* out, sma_stream_t s) {
  if (!x || !out || B <= 0 || H <= 0 || W <= 0 || C <= 0 || ld < C) return SMA_ERR_BAD_ARG;
  to_uint8_kernel<<<nblocks((long long)B * H * W * C), 256, 0, as_stream(s)>>>(x, (long long)B * H * W, C, ld, bgr, out);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_nchw_to_nhwc(const float* x, int B, int C, int H, int W, float* y, int yld, sma_stream_t s) {
  if (!x || !y || B <= 0 || C <= 0 || H <= 0 || W <= 0 || yld < C) return SMA_ERR_BAD_ARG;
  nchw_to_nhwc_kernel<<<nblocks((long long)B * C * H * W), 256, 0, as_stream(s)>>>(x, B, C, H * W, y, yld);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
extern "C" int sma_nhwc_to_nchw(const float* x, int B, int C, int H, int W, int xld, float* y, sma_stream_t s) {
  if (!x || !y || B <= 0 || C <= 0 || H <= 0 || W <= 0 || xld < C) return SMA_ERR_BAD_ARG;
  nhwc_to_nchw_kernel<<<nblocks((long long)B * C * H * W), 256, 0, as_stream(s)>>>(x, B, C, H * W, xld, y);
  SMA_LAUNCH_CHECK(); return SMA_OK;
}
