// GroupNorm statistics (-> per-(b,c) scale/shift for the conv prologue), affine+activation, LayerNorm.
#include "sma_common.cuh"
#include <cstdlib>

namespace {

constexpr int GN_CHUNK = 256;   // pixels per partial block

// partial[(b*nchunk+chunk)*C*2 + c*2 + {0,1}] = sum, sum of squares over the chunk's pixels for channel c
__global__ void gn_partial_kernel(const float* __restrict__ x, int HW, int C, long long bstride, int ld, float* __restrict__ partial, int nchunk) {
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int p0 = chunk * GN_CHUNK, p1 = min(HW, p0 + GN_CHUNK);
  extern __shared__ float sm[];          // [2][rows][C] reduction scratch
  const int tid = threadIdx.x;           // 256 threads: lane -> channel (coalesced), row groups -> pixels
  const int cpt = min(C, 256);           // channels covered per pass
  const int rows = 256 / cpt;            // pixel rows processed concurrently
  const int c_in = tid % cpt, r_in = tid / cpt;
  for (int cb = 0; cb < C; cb += cpt) {
    float s = 0.f, q = 0.f;
    int c = cb + c_in;
    if (r_in < rows && c < C) {
      const float* px = x + (long long)b * bstride + c;
      for (int p = p0 + r_in; p < p1; p += rows) { float v = __ldg(px + (long long)p * ld); s += v; q = fmaf(v, v, q); }
    }
    sm[tid] = s; sm[256 + tid] = q;
    __syncthreads();
    if (tid < cpt && cb + tid < C) {
      float ss = 0.f, qq = 0.f;
      for (int r = 0; r < rows; r++) { ss += sm[r * cpt + tid]; qq += sm[256 + r * cpt + tid]; }
      float* o = partial + (((long long)b * nchunk + chunk) * C + cb + tid) * 2;
      o[0] = ss; o[1] = qq;
    }
    __syncthreads();
  }
}

// same partial sums with 128-bit loads: a thread owns 4 consecutive channels and every `rows`-th pixel of the chunk, 4 loads in flight
__global__ void gn_partial4_kernel(const float* __restrict__ x, int HW, int C, long long bstride, int ld, float* __restrict__ partial, int nchunk) {
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int p0 = chunk * GN_CHUNK, p1 = min(HW, p0 + GN_CHUNK);
  __shared__ float sm[8][256];
  const int tid = threadIdx.x;
  const int nq = C >> 2;                 // channel quads per pixel
  const int cpt = min(nq, 256);          // quads covered per pass
  const int rows = 256 / cpt;            // pixels processed concurrently
  const int c_in = tid % cpt, r_in = tid / cpt;
  const int ld4 = ld >> 2;
  for (int cb = 0; cb < nq; cb += cpt) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    const int cq = cb + c_in;
    if (r_in < rows && cq < nq) {
      const float4* px = reinterpret_cast<const float4*>(x + (long long)b * bstride) + cq;
      int p = p0 + r_in;
      for (; p + 3 * rows < p1; p += 4 * rows) {
        const float4 v0 = __ldg(px + (long long)p * ld4), v1 = __ldg(px + (long long)(p + rows) * ld4), v2 = __ldg(px + (long long)(p + 2 * rows) * ld4),
                     v3 = __ldg(px + (long long)(p + 3 * rows) * ld4);
        s.x += (v0.x + v1.x) + (v2.x + v3.x); s.y += (v0.y + v1.y) + (v2.y + v3.y); s.z += (v0.z + v1.z) + (v2.z + v3.z); s.w += (v0.w + v1.w) + (v2.w + v3.w);
        q.x = fmaf(v0.x, v0.x, fmaf(v1.x, v1.x, fmaf(v2.x, v2.x, fmaf(v3.x, v3.x, q.x)))); q.y = fmaf(v0.y, v0.y, fmaf(v1.y, v1.y, fmaf(v2.y, v2.y, fmaf(v3.y, v3.y, q.y))));
        q.z = fmaf(v0.z, v0.z, fmaf(v1.z, v1.z, fmaf(v2.z, v2.z, fmaf(v3.z, v3.z, q.z)))); q.w = fmaf(v0.w, v0.w, fmaf(v1.w, v1.w, fmaf(v2.w, v2.w, fmaf(v3.w, v3.w, q.w))));
      }
      for (; p < p1; p += rows) {
        const float4 v = __ldg(px + (long long)p * ld4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
      }
    }
    sm[0][tid] = s.x; sm[1][tid] = q.x; sm[2][tid] = s.y; sm[3][tid] = q.y; sm[4][tid] = s.z; sm[5][tid] = q.z; sm[6][tid] = s.w; sm[7][tid] = q.w;
    __syncthreads();
    if (tid < cpt && cb + tid < nq) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int r = 0; r < rows; r++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] += sm[i][r * cpt + tid];
      }
      float4* o = reinterpret_cast<float4*>(partial + (((long long)b * nchunk + chunk) * C + (cb + tid) * 4) * 2);     // [c][{sum, sumsq}] x 4 channels
      o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]); o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncthreads();
  }
}

// one block per (b, group): combine chunk partials in fp64, emit scale/shift per channel.  `per` = channels per partial entry: 1 (the standalone
// statistics pass) or 2 (channel pairs: the partial sums a convolution's epilogue produced, conv_tc.cu gn_pairs_reduce_store).  `fold` > 1: the
// producer was a depth-to-space (un-patchify) convolution whose `fold` sub-pixel column blocks of C channels each all belong to the same C output
// channels (partial row = fold * C / per entries; HW counts the PRODUCER's rows, so a group holds HW * fold * cg values).  Output rows have pitch out_ld.
__global__ void gn_finalize_kernel(const float* __restrict__ partial, int nchunk, int C, int groups, int HW, float eps,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ scale, float* __restrict__ shift, int per,
                                   int fold, int out_ld) {
  const int g = blockIdx.x, b = blockIdx.y;
  const int cg = C / groups, eg = cg / per, Cp = C / per, rowp = Cp * fold;
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < nchunk * fold * eg; i += blockDim.x) {
    const int e = i % eg; const int t = i / eg; const int f = t % fold, chunk = t / fold;
    const float* o = partial + (((long long)b * nchunk + chunk) * rowp + f * Cp + g * eg + e) * 2;
    s += (double)o[0]; q += (double)o[1];
  }
  __shared__ double sh[2][32];
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ss = 0, qq = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) { ss += sh[0][i]; qq += sh[1][i]; }
    double n = (double)HW * fold * cg; double mean = ss / n; double var = qq / n - mean * mean; if (var < 0) var = 0;
    sh[0][0] = mean; sh[1][0] = 1.0 / sqrt(var + (double)eps);
  }
  __syncthreads();
  float mean = (float)sh[0][0], rstd = (float)sh[1][0];
  for (int i = threadIdx.x; i < cg; i += blockDim.x) {
    int c = g * cg + i;
    float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    float sc = ga * rstd;
    scale[(long long)b * out_ld + c] = sc; shift[(long long)b * out_ld + c] = be - mean * sc;
  }
}

__global__ void affine_act_kernel(const float* __restrict__ x, int HW, int C, long long bstride, int ld, const float* __restrict__ scale,
                                  const float* __restrict__ shift, int act, float* __restrict__ y, long long ybs, int yld, long long total4) {
  const int C4 = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4); long long pp = i / C4; int p = (int)(pp % HW); int b = (int)(pp / HW);
    float4 v = __ldg(reinterpret_cast<const float4*>(x + (long long)b * bstride + (long long)p * ld + c4 * 4));
    float4 s = make_float4(1.f, 1.f, 1.f, 1.f), h = make_float4(0.f, 0.f, 0.f, 0.f);
    if (scale) { s = __ldg(reinterpret_cast<const float4*>(scale + (long long)b * C + c4 * 4)); h = __ldg(reinterpret_cast<const float4*>(shift + (long long)b * C + c4 * 4)); }
    v.x = sma_act(fmaf(v.x, s.x, h.x), act); v.y = sma_act(fmaf(v.y, s.y, h.y), act);
    v.z = sma_act(fmaf(v.z, s.z, h.z), act); v.w = sma_act(fmaf(v.w, s.w, h.w), act);
    *reinterpret_cast<float4*>(y + (long long)b * ybs + (long long)p * yld + c4 * 4) = v;
  }
}

// one warp per token row; E in {32,...,1024}, E % 32 == 0; two-pass mean/variance in registers (E = 256 on the product path; E = 32 runs on layernorm32_kernel below)
template <int EPL>   // elements per lane
__global__ void layernorm_kernel(const float* __restrict__ x, int rows, const float* __restrict__ g, const float* __restrict__ be, float eps,
                                 const float* __restrict__ pos, int pos_rows, float* __restrict__ y, float* __restrict__ yq) {
  constexpr int E = EPL * 32;
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float v[EPL]; float s = 0.f;
#pragma unroll
  for (int i = 0; i < EPL; i++) { v[i] = __ldg(x + (long long)row * E + lane + i * 32); s += v[i]; }
  float mean = warp_sum(s) / E; float q = 0.f;
#pragma unroll
  for (int i = 0; i < EPL; i++) { float d = v[i] - mean; q = fmaf(d, d, q); }
  float rstd = 1.f / sqrtf(warp_sum(q) / E + eps);
#pragma unroll
  for (int i = 0; i < EPL; i++) {
    int c = lane + i * 32;
    float o = (v[i] - mean) * rstd * __ldg(g + c) + __ldg(be + c);
    if (y) y[(long long)row * E + c] = o;
    if (yq) yq[(long long)row * E + c] = o + __ldg(pos + (long long)(row % pos_rows) * E + c);
  }
}

// E = 32 (the motion transformer): eight lanes per token row with one 128-bit load each, two rows per thread in flight (the warp-per-row form above keeps 4 bytes per
// thread in flight and runs at 0.7 TB/s on the 65 536-row tensors of a 64-frame step).  Two-pass mean / variance like the general kernel.
__global__ void __launch_bounds__(256) layernorm32_kernel(const float* __restrict__ x, int rows, const float* __restrict__ g, const float* __restrict__ be, float eps,
                                                          const float* __restrict__ pos, int pos_rows, float* __restrict__ y, float* __restrict__ yq) {
  const int sub = threadIdx.x & 7;
  const long long r0 = (long long)blockIdx.x * 64 + (threadIdx.x >> 3);
  float4 v[2];
#pragma unroll
  for (int u = 0; u < 2; u++) {
    const long long row = r0 + u * 32;
    v[u] = row < rows ? __ldg(reinterpret_cast<const float4*>(x + row * 32) + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + sub), b4 = __ldg(reinterpret_cast<const float4*>(be) + sub);
#pragma unroll
  for (int u = 0; u < 2; u++) {
    const long long row = r0 + u * 32;
    float s = (v[u].x + v[u].y) + (v[u].z + v[u].w);
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
    const float mean = s * (1.f / 32.f);
    const float dx = v[u].x - mean, dy = v[u].y - mean, dz = v[u].z - mean, dw = v[u].w - mean;
    float q = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
    q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 4);
    const float rstd = 1.f / sqrtf(q * (1.f / 32.f) + eps);
    if (row >= rows) continue;
    const float4 o = make_float4(dx * rstd * g4.x + b4.x, dy * rstd * g4.y + b4.y, dz * rstd * g4.z + b4.z, dw * rstd * g4.w + b4.w);
    if (y) *(reinterpret_cast<float4*>(y + row * 32) + sub) = o;
    if (yq) {
      const float4 p4 = __ldg(reinterpret_cast<const float4*>(pos + (row % pos_rows) * 32) + sub);
      *(reinterpret_cast<float4*>(yq + row * 32) + sub) = make_float4(o.x + p4.x, o.y + p4.y, o.z + p4.z, o.w + p4.w);
    }
  }
}

}  // namespace

extern "C" int sma_groupnorm_stats(const float* x, int B, int HW, int C, int64_t bstride, int ld, int groups, float eps, const float* gamma,
                                   const float* beta, float* partial, float* scale, float* shift, sma_stream_t stream) {
  if (!x || !partial || !scale || !shift || B <= 0 || HW <= 0 || C <= 0 || groups <= 0 || C % groups) return SMA_ERR_BAD_ARG;
  int nchunk = cdiv(HW, GN_CHUNK);
  static const bool force_scalar = getenv("SMA_GN_SCALAR") != nullptr;     // debugging aid
  const bool vec = !force_scalar && (C & 3) == 0 && (ld & 3) == 0 && (bstride & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(partial)) & 15) == 0;
  if (vec) gn_partial4_kernel<<<dim3(nchunk, B), 256, 0, as_stream(stream)>>>(x, HW, C, bstride, ld, partial, nchunk);
  else gn_partial_kernel<<<dim3(nchunk, B), 256, 512 * sizeof(float), as_stream(stream)>>>(x, HW, C, bstride, ld, partial, nchunk);
  SMA_LAUNCH_CHECK();
  gn_finalize_kernel<<<dim3(groups, B), 128, 0, as_stream(stream)>>>(partial, nchunk, C, groups, HW, eps, gamma, beta, scale, shift, 1, 1, C);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}

extern "C" int sma_groupnorm_finalize_pairs(const float* partial, int B, int nchunk, int C, int groups, int HW, int fold, float eps, const float* gamma,
                                            const float* beta, float* scale, float* shift, int out_ld, sma_stream_t stream) {
  if (!partial || !scale || !shift || B <= 0 || nchunk <= 0 || HW <= 0 || C <= 0 || groups <= 0 || C % groups || ((C / groups) & 1) || fold < 1 || out_ld < C) return SMA_ERR_BAD_ARG;
  gn_finalize_kernel<<<dim3(groups, B), 256, 0, as_stream(stream)>>>(partial, nchunk, C, groups, HW, eps, gamma, beta, scale, shift, 2, fold, out_ld);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}

extern "C" int sma_affine_act(const float* x, int B, int HW, int C, int64_t bstride, int ld, const float* scale, const float* shift, int act,
                              float* y, int64_t ybs, int yld, sma_stream_t stream) {
  if (!x || !y || B <= 0 || HW <= 0 || C <= 0) return SMA_ERR_BAD_ARG;
  if ((C & 3) || (ld & 3) || (yld & 3) || (bstride & 3) || (ybs & 3)) return SMA_ERR_UNSUPPORTED;
  if ((scale == nullptr) != (shift == nullptr)) return SMA_ERR_BAD_ARG;
  long long total4 = (long long)B * HW * (C >> 2);
  int blocks = (int)((total4 + 255) / 256); if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  affine_act_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, HW, C, bstride, ld, scale, shift, act, y, ybs, yld, total4);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}

extern "C" int sma_layernorm(const float* x, int rows, int E, const float* gamma, const float* beta, float eps, const float* pos, int pos_rows,
                             float* y, float* yq, sma_stream_t stream) {
  if (!x || !gamma || !beta || rows <= 0 || (!y && !yq) || (yq && (!pos || pos_rows <= 0))) return SMA_ERR_BAD_ARG;
  dim3 grid(cdiv(rows, 8));
  cudaStream_t st = as_stream(stream);
  if (!pos_rows) pos_rows = 1;
  const bool al16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(pos) |
                      reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(yq)) & 15) == 0;
  if (E == 32 && al16) layernorm32_kernel<<<cdiv(rows, 64), 256, 0, st>>>(x, rows, gamma, beta, eps, pos, pos_rows, y, yq);
  else if (E == 32) layernorm_kernel<1><<<grid, 256, 0, st>>>(x, rows, gamma, beta, eps, pos, pos_rows, y, yq);
  else if (E == 256) layernorm_kernel<8><<<grid, 256, 0, st>>>(x, rows, gamma, beta, eps, pos, pos_rows, y, yq);
  else return SMA_ERR_UNSUPPORTED;
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}
