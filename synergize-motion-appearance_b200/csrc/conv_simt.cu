// CUDA-core (FFMA) implicit-GEMM convolution, NHWC fp32.
//
// This is the exact-fp32 kernel of sma_conv2d_fwd.  It serves every conv/linear whose shape does not
// fit the tcgen05 3xTF32 kernel in conv_tc.cu (Cin in {2,3,15,35}, 7x7 heads, Cout in {1,2,3,17,75},
// tiny M) and is the numerical yard-stick the tensor-core kernel is tested against on the GPU.
//
// GEMM view: M = B*Ho*Wo output pixels, N = Cout, K = kh*kw*Cin with k = (ky*kw+kx)*Cin + c so that a
// K-chunk is contiguous in NHWC memory.  256 threads, BMxBN tile, BK = 16, register tile TMxTN,
// double-buffered shared memory with the global loads of tile t+1 in flight during the FMAs of tile t.
// Fused prologue (GroupNorm-apply + swish on the operand load, nearest x2 upsample, zero padding after
// the normalisation) and epilogue (bias, activation, residual add, depth-to-space / NCHW store).
#include "sma_common.cuh"

namespace {

constexpr int BK = 16;

struct ConvP {
  const float* x; const float* w; const float* bias; const float* pre_scale; const float* pre_shift;
  const float* res; float* y;
  long long in_bs, out_bs, res_bs;
  int B, Hi, Wi, Cin, in_ld, ldw, Cout, kh, kw, stride, pad_t, pad_l, up, pre_act;
  int Ho, Wo, out_ld, act, res_ld, d2s, out_nchw;
  int M, K, HoWo, ncol4;
};

template <int BM, int BN, int TM, int TN, bool VEC>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvP p) {
  constexpr int NTX = BN / TN;           // threads along N
  constexpr int RG = TM / 4, CG = TN / 4;
  constexpr int AF4 = BM / 64;           // float4 A loads per thread (VEC)
  constexpr int BF4 = (BN >= 64) ? BN / 64 : 1;
  constexpr int AS = BM + 4;
  __shared__ __align__(16) float As[2][BK][AS];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid % NTX, ty = tid / NTX;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int Hv = p.Hi << p.up, Wv = p.Wi << p.up;

  // ---- per-thread A-row bookkeeping (VEC path): rows tid/4 + i*64, k-quad tid%4 ----
  const float* a_base[AF4]; int a_iy0[AF4], a_ix0[AF4], a_b[AF4]; bool a_ok[AF4];
  const int kq = tid & 3;
  if (VEC) {
#pragma unroll
    for (int i = 0; i < AF4; i++) {
      int m = m0 + (tid >> 2) + i * 64;
      a_ok[i] = m < p.M;
      int mm = a_ok[i] ? m : 0;
      int b = mm / p.HoWo; int r = mm - b * p.HoWo; int oy = r / p.Wo; int ox = r - oy * p.Wo;
      a_b[i] = b; a_base[i] = p.x + (long long)b * p.in_bs;
      a_iy0[i] = oy * p.stride - p.pad_t; a_ix0[i] = ox * p.stride - p.pad_l;
    }
  }
  int c0 = 0, ky = 0, kx = 0;            // running decode of the K-chunk start (VEC path)

  float4 ra[AF4 > 0 ? AF4 : 1]; float rs[VEC ? 1 : BM / 16];
  float4 rb[BF4];

  auto load_tile = [&](int k0) {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < AF4; i++) {
        int iy = a_iy0[i] + ky, ix = a_ix0[i] + kx;
        bool inb = a_ok[i] && (unsigned)iy < (unsigned)Hv && (unsigned)ix < (unsigned)Wv;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inb) {
          int c = c0 + kq * 4;
          v = __ldg(reinterpret_cast<const float4*>(a_base[i] + ((long long)(iy >> p.up) * p.Wi + (ix >> p.up)) * p.in_ld + c));
          if (p.pre_scale) {
            float4 s = __ldg(reinterpret_cast<const float4*>(p.pre_scale + (long long)a_b[i] * p.Cin + c));
            float4 h = __ldg(reinterpret_cast<const float4*>(p.pre_shift + (long long)a_b[i] * p.Cin + c));
            v.x = sma_act(fmaf(v.x, s.x, h.x), p.pre_act); v.y = sma_act(fmaf(v.y, s.y, h.y), p.pre_act);
            v.z = sma_act(fmaf(v.z, s.z, h.z), p.pre_act); v.w = sma_act(fmaf(v.w, s.w, h.w), p.pre_act);
          }
        }
        ra[i] = v;
      }
      c0 += BK; if (c0 >= p.Cin) { c0 = 0; if (++kx == p.kw) { kx = 0; ++ky; } }
    } else {
#pragma unroll
      for (int i = 0; i < BM / 16; i++) {
        int e = tid + i * 256; int row = e >> 4, kk = e & 15;
        int m = m0 + row, k = k0 + kk; float v = 0.f;
        if (m < p.M && k < p.K) {
          int b = m / p.HoWo; int r = m - b * p.HoWo; int oy = r / p.Wo; int ox = r - oy * p.Wo;
          int tap = k / p.Cin; int c = k - tap * p.Cin; int kyy = tap / p.kw; int kxx = tap - kyy * p.kw;
          int iy = oy * p.stride - p.pad_t + kyy, ix = ox * p.stride - p.pad_l + kxx;
          if ((unsigned)iy < (unsigned)Hv && (unsigned)ix < (unsigned)Wv) {
            v = __ldg(p.x + (long long)b * p.in_bs + ((long long)(iy >> p.up) * p.Wi + (ix >> p.up)) * p.in_ld + c);
            if (p.pre_scale) v = sma_act(fmaf(v, __ldg(p.pre_scale + (long long)b * p.Cin + c), __ldg(p.pre_shift + (long long)b * p.Cin + c)), p.pre_act);
          }
        }
        rs[i] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < BF4; i++) {
      int f = tid + i * 256; int row = f / (BN / 4), c4 = f % (BN / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      int k = k0 + row, n = n0 + c4 * 4;
      if (row < BK && k < p.K && n < p.ncol4) v = __ldg(reinterpret_cast<const float4*>(p.w + (long long)k * p.ldw + n));
      rb[i] = v;
    }
  };
  auto store_tile = [&](int buf) {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < AF4; i++) {
        int row = (tid >> 2) + i * 64;
        As[buf][kq * 4 + 0][row] = ra[i].x; As[buf][kq * 4 + 1][row] = ra[i].y;
        As[buf][kq * 4 + 2][row] = ra[i].z; As[buf][kq * 4 + 3][row] = ra[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < BM / 16; i++) { int e = tid + i * 256; As[buf][e & 15][e >> 4] = rs[i]; }
    }
#pragma unroll
    for (int i = 0; i < BF4; i++) {
      int f = tid + i * 256; int row = f / (BN / 4), c4 = f % (BN / 4);
      if (row < BK) *reinterpret_cast<float4*>(&Bs[buf][row][c4 * 4]) = rb[i];
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  const int nt = (p.K + BK - 1) / BK;
  load_tile(0); store_tile(0); __syncthreads();
  for (int t = 0; t < nt; t++) {
    const int buf = t & 1;
    if (t + 1 < nt) load_tile((t + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < RG; g++) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][k][g * (BM / RG) + ty * 4]);
        a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < CG; g++) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][g * (BN / CG) + tx * 4]);
        b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < nt) store_tile(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue ----
  const bool vec_out = !p.out_nchw && p.d2s <= 1 && (p.out_ld & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                       (p.out_bs & 3) == 0 && (p.Cout & 3) == 0 &&
                       (!p.res || ((p.res_ld & 3) == 0 && (p.res_bs & 3) == 0 && (reinterpret_cast<uintptr_t>(p.res) & 15) == 0));
#pragma unroll
  for (int i = 0; i < TM; i++) {
    int row = (i / 4) * (BM / RG) + ty * 4 + (i & 3);
    int m = m0 + row;
    if (m >= p.M) continue;
    int b = m / p.HoWo; int r = m - b * p.HoWo; int oy = r / p.Wo; int ox = r - oy * p.Wo;
#pragma unroll
    for (int g = 0; g < CG; g++) {
      int n = n0 + g * (BN / CG) + tx * 4;
      if (n >= p.Cout) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float t = acc[i][g * 4 + j];
        if (p.bias && n + j < p.Cout) t += __ldg(p.bias + n + j);
        v[j] = sma_act(t, p.act);
      }
      if (vec_out) {
        if (p.res) {
          float4 rr = __ldg(reinterpret_cast<const float4*>(p.res + (long long)b * p.res_bs + (long long)r * p.res_ld + n));
          v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
        }
        *reinterpret_cast<float4*>(p.y + (long long)b * p.out_bs + (long long)r * p.out_ld + n) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          int nn = n + j;
          if (nn >= p.Cout) break;
          float t = v[j];
          if (p.res) t += __ldg(p.res + (long long)b * p.res_bs + (long long)r * p.res_ld + nn);
          if (p.out_nchw) {
            p.y[(((long long)b * p.Cout + nn) * p.Ho + oy) * p.Wo + ox] = t;
          } else if (p.d2s > 1) {
            int C = p.Cout / (p.d2s * p.d2s); int q = nn / C; int c = nn - q * C; int p1 = q / p.d2s; int p2 = q - p1 * p.d2s;
            long long pix = (long long)(oy * p.d2s + p1) * (p.Wo * p.d2s) + (ox * p.d2s + p2);
            p.y[(long long)b * p.out_bs + pix * p.out_ld + c] = t;
          } else {
            p.y[(long long)b * p.out_bs + (long long)r * p.out_ld + nn] = t;
          }
        }
      }
    }
  }
}

template <int BM, int BN, int TM, int TN>
int launch(const ConvP& p, bool vec, cudaStream_t st) {
  dim3 grid(cdiv(p.M, BM), cdiv(p.Cout, BN));
  if (vec) conv_simt_kernel<BM, BN, TM, TN, true><<<grid, 256, 0, st>>>(p);
  else conv_simt_kernel<BM, BN, TM, TN, false><<<grid, 256, 0, st>>>(p);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}

__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ bias, int Cout, int Cin, int kh, int kw,
                                   const float* g, const float* be, const float* mu, const float* var, float eps,
                                   float* __restrict__ wp, int ldw, float* __restrict__ bo) {
  long long K = (long long)kh * kw * Cin;
  long long total = K * ldw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int n = (int)(i % ldw); long long k = i / ldw;
    float v = 0.f;
    if (n < Cout) {
      int c = (int)(k % Cin); int tap = (int)(k / Cin); int ky = tap / kw, kx = tap % kw;
      v = w[(((long long)n * Cin + c) * kh + ky) * kw + kx];
      if (g) v *= g[n] / sqrtf(var[n] + eps);
    }
    wp[i] = v;
  }
  if (bo && blockIdx.x == 0) {
    for (int n = threadIdx.x; n < Cout; n += blockDim.x) {
      float b = bias ? bias[n] : 0.f;
      if (g) b = (b - mu[n]) * (g[n] / sqrtf(var[n] + eps)) + be[n];
      bo[n] = b;
    }
  }
}

}  // namespace

int sma_conv2d_tc_try(sma_conv_desc* d, cudaStream_t st);   // conv_tc.cu ; returns SMA_ERR_UNSUPPORTED when not applicable
int sma_conv2d_ts_try(sma_conv_desc* d, cudaStream_t st);   // conv_ts.cu ; likewise

extern "C" int sma_conv2d_fwd(sma_conv_desc* d, sma_stream_t stream) {
  if (!d || !d->x || !d->w || !d->y) return SMA_ERR_BAD_ARG;
  if (d->B <= 0 || d->Hi <= 0 || d->Wi <= 0 || d->Cin <= 0 || d->Cout <= 0 || d->kh <= 0 || d->kw <= 0 || d->stride <= 0 ||
      d->Ho <= 0 || d->Wo <= 0 || d->ldw < d->Cout || (d->ldw & 3))
    return SMA_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(d->w) & 15)) return SMA_ERR_BAD_ARG;
  if ((d->pre_scale == nullptr) != (d->pre_shift == nullptr)) return SMA_ERR_BAD_ARG;
  if (d->d2s > 1 && (d->res || d->out_nchw || d->Cout % (d->d2s * d->d2s))) return SMA_ERR_UNSUPPORTED;
  if ((long long)d->B * d->Ho * d->Wo > 0x7fffffffLL) return SMA_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  d->gn_chunks = 0;                                   // set by the persistent tensor-core kernel when it produces the fused GroupNorm partial sums
  if (d->precision != SMA_PREC_EXACT) {
    int r = sma_conv2d_ts_try(d, st);
    if (r != SMA_ERR_UNSUPPORTED) { d->kernel_used = 4; return r; }
    r = sma_conv2d_tc_try(d, st);
    if (r != SMA_ERR_UNSUPPORTED) return r;
  }
  d->kernel_used = 0;
  if (d->plan_only) return SMA_OK;
  if (d->aux || d->x2 || d->split_ws) return SMA_ERR_UNSUPPORTED;           // the SFT epilogue / the two-tensor input live in the persistent tensor-core kernel only: callers fall back to sma_sft_combine
  ConvP p;
  p.x = d->x; p.w = d->w; p.bias = d->bias; p.pre_scale = d->pre_scale; p.pre_shift = d->pre_shift; p.res = d->res; p.y = d->y;
  p.in_bs = d->in_bstride; p.out_bs = d->out_bstride; p.res_bs = d->res_bstride;
  p.B = d->B; p.Hi = d->Hi; p.Wi = d->Wi; p.Cin = d->Cin; p.in_ld = d->in_ld; p.ldw = d->ldw; p.Cout = d->Cout;
  p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.up = d->upsample2 ? 1 : 0;
  p.pre_act = d->pre_act; p.Ho = d->Ho; p.Wo = d->Wo; p.out_ld = d->out_ld; p.act = d->act; p.res_ld = d->res_ld;
  p.d2s = d->d2s; p.out_nchw = d->out_nchw;
  p.HoWo = d->Ho * d->Wo; p.M = d->B * p.HoWo; p.K = d->kh * d->kw * d->Cin; p.ncol4 = (d->Cout + 3) & ~3;
  const bool vec = (d->Cin % BK == 0) && (d->in_ld % 4 == 0) && (d->in_bstride % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(d->x) & 15) == 0) &&
                   (!d->pre_scale || (((reinterpret_cast<uintptr_t>(d->pre_scale) | reinterpret_cast<uintptr_t>(d->pre_shift)) & 15) == 0));
  const int N = d->Cout;
  const long long big_tiles = (long long)cdiv(p.M, 128) * cdiv(N, N <= 32 ? 32 : (N <= 64 ? 64 : 128));
  if (N <= 32) return launch<128, 32, 4, 4>(p, vec, st);
  if (big_tiles < 2 * kNumSMs) {
    if (N <= 64) return launch<64, 64, 4, 4>(p, vec, st);
    return launch<64, 128, 4, 8>(p, vec, st);
  }
  if (N <= 64) return launch<128, 64, 8, 4>(p, vec, st);
  return launch<128, 128, 8, 8>(p, vec, st);
}

extern "C" int sma_pack_conv_weight(const float* w, const float* bias, int Cout, int Cin, int kh, int kw, const float* g,
                                    const float* be, const float* mu, const float* var, float eps, float* wp, int ldw,
                                    float* bo, sma_stream_t stream) {
  if (!w || !wp || Cout <= 0 || Cin <= 0 || kh <= 0 || kw <= 0 || ldw < Cout || (ldw & 3)) return SMA_ERR_BAD_ARG;
  if (g && (!be || !mu || !var)) return SMA_ERR_BAD_ARG;
  long long total = (long long)kh * kw * Cin * ldw;
  int blocks = (int)((total + 255) / 256); if (blocks > 4096) blocks = 4096;
  pack_weight_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w, bias, Cout, Cin, kh, kw, g, be, mu, var, eps, wp, ldw, bo);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}
