// tcgen05 implicit-GEMM convolution for the update block (v2 of conv_ffma.cu): A and B tiles staged by TMA into
// 128B-swizzled shared memory, tcgen05.mma (kind::f16, M128 x N<=256 x K16) issued by one thread, fp32 accumulators in
// TMEM, epilogue warps read them back with tcgen05.ld and apply the update block's fusions.
// Replaces the same reference code as conv_ffma.cu: core/update.py:33-60, 79-97, 6-14, 123-126 and the 3x3 layers of
// core/interp_weights_est.py:10-47.
//
// fp32-faithful on fp16 tensor cores: every activation x and weight w is carried as an exact-sum pair of halves
// (x = x_hi + x_lo, |x_lo| <= ulp(x_hi)/2; weights pre-scaled by a power of two so w_lo stays normal) and the product is
// accumulated as  x_hi*w_hi + x_hi*w_lo + x_lo*w_hi  — 3 MMAs per K step, relative error ~2^-21 per term, which keeps the
// 32-iteration recurrence inside the 1e-3 EPE budget where plain TF32/bf16 operands do not.
//
// GEMM view: M = 128 pixels (a TH x TW patch of one image), N = Cout tile, K = taps x channels in blocks of 64.
// Zero padding comes for free: the A tile of filter tap (dy,dx) is a 4-D TMA box at (c, x0+dx, y0+dy, b) and
// out-of-image elements are zero-filled by the TMA unit.
#include "umma_ptx.cuh"

namespace rnc {
namespace umma {

constexpr int kThreads = 192;        // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int kBM = 128;             // pixels per tile
constexpr int kBK = 64;              // channels per stage (64 halves = one 128-byte swizzle row)
constexpr int kATile = kBM * kBK * 2;   // 16 KB per half-plane

struct Params {
  // tile geometry
  int B, H, W, TW, TH, tiles_x, tiles_y;
  int kw, ph, pw, ntaps;
  int nblk0, nblk;                   // 64-channel blocks in segment 0 / in total (per tap)
  int cout, epilogue;
  float unscale;
  const float* bias;
  float* out_f32; int ldo_f32;
  __half* out_hi; __half* out_lo; int ldo_split;
  float* h; int ldh;
  float* aux0; int ldaux;
};

// exact hi/lo split of 8 floats into two 16-byte vectors of halves
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = fminf(fmaxf(v[2 * i], -65504.f), 65504.f), b = fminf(fmaxf(v[2 * i + 1], -65504.f), 65504.f);
    const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
    h[i] = __halves2half2(ha, hb);
    l[i] = __halves2half2(__float2half_rn(a - __half2float(ha)), __float2half_rn(b - __half2float(hb)));
  }
  hi = *reinterpret_cast<uint4*>(h);
  lo = *reinterpret_cast<uint4*>(l);
}

template <int BN>
struct Cfg {
  static constexpr int kBTile = BN * kBK * 2;                     // bytes per half-plane of weights
  static constexpr int kStage = 2 * kATile + 2 * kBTile;
  static constexpr int kStages = (200 * 1024) / kStage > 4 ? 4 : (200 * 1024) / kStage;
  static constexpr int kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int kSmem = kStages * kStage + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap mA0h, const __grid_constant__ CUtensorMap mA0l,
                 const __grid_constant__ CUtensorMap mA1h, const __grid_constant__ CUtensorMap mA1l,
                 const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, const Params p) {
  using C = Cfg<BN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStage);
  uint64_t* empty = full + C::kStages;
  uint64_t* tmem_full = empty + C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int tpi = p.tiles_x * p.tiles_y;
  const int b = tile / tpi, tr = tile - b * tpi;
  const int y0 = (tr / p.tiles_x) * p.TH, x0 = (tr % p.tiles_x) * p.TW;
  const int n0 = blockIdx.y * BN;
  const int nk = p.ntaps * p.nblk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int it = 0; it < nk; ++it) {
        const int s = it % C::kStages, ph = (it / C::kStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        unsigned char* st = smem + s * C::kStage;
        mbar_expect_tx(&full[s], C::kStage);
        const int tap = it / p.nblk, cb = it - tap * p.nblk;
        const int dy = tap / p.kw - p.ph, dx = tap % p.kw - p.pw;
        const bool seg0 = cb < p.nblk0;
        const int c = (seg0 ? cb : cb - p.nblk0) * kBK;
        tma_load_4d(st, seg0 ? &mA0h : &mA1h, &full[s], c, x0 + dx, y0 + dy, b);
        tma_load_4d(st + kATile, seg0 ? &mA0l : &mA1l, &full[s], c, x0 + dx, y0 + dy, b);
        tma_load_2d(st + 2 * kATile, &mBh, &full[s], it * kBK, n0);
        tma_load_2d(st + 2 * kATile + C::kBTile, &mBl, &full[s], it * kBK, n0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6), A=B=F16, K-major, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(kBM >> 4) << 24);
      for (int it = 0; it < nk; ++it) {
        const int s = it % C::kStages, ph = (it / C::kStages) & 1;
        mbar_wait(&full[s], ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + s * C::kStage);
        const uint64_t ah = smem_desc_sw128(sa), al = smem_desc_sw128(sa + kATile);
        const uint64_t bh = smem_desc_sw128(sa + 2 * kATile), bl = smem_desc_sw128(sa + 2 * kATile + C::kBTile);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          const uint64_t ko = static_cast<uint64_t>((k * 32) >> 4);     // advance 16 halves = 32 B inside the swizzle row
          umma_f16(tmem_base, ah + ko, bh + ko, idesc, (it | k) != 0);
          umma_f16(tmem_base, ah + ko, bl + ko, idesc, 1);
          umma_f16(tmem_base, al + ko, bh + ko, idesc, 1);
        }
        umma_commit(&empty[s]);          // frees the stage once these MMAs have read it
      }
      umma_commit(tmem_full);            // accumulator complete
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> registers -> global
    const int lg = warp & 3;                       // TMEM lane group this warp may access
    const int ml = lg * 32 + lane;                 // row of the tile = pixel
    const int y = y0 + ml / p.TW, x = x0 + ml % p.TW;
    const bool valid = y < p.H && x < p.W;
    const size_t pix = (static_cast<size_t>(b) * p.H + y) * p.W + x;
    mbar_wait(tmem_full, 0);
    tcgen05_fence_after();
    const int epi = p.epilogue;
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; ++cc) {
      const int n = n0 + cc * 32;
      if (n >= p.cout + (epi == RNC_EPI_RELU_FLOW ? 2 : 0)) break;   // warp-uniform
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + cc * 32, r);
      if (!valid) continue;
      float v[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n) + q);
        v[4 * q + 0] = fmaf(__uint_as_float(r[4 * q + 0]), p.unscale, bv.x);
        v[4 * q + 1] = fmaf(__uint_as_float(r[4 * q + 1]), p.unscale, bv.y);
        v[4 * q + 2] = fmaf(__uint_as_float(r[4 * q + 2]), p.unscale, bv.z);
        v[4 * q + 3] = fmaf(__uint_as_float(r[4 * q + 3]), p.unscale, bv.w);
      }
      if (epi == RNC_EPI_GRU_ZR) {
        const int Ch = p.cout >> 1;
        if (n < Ch) {            // z gate -> fp32 aux buffer
          float4* dst = reinterpret_cast<float4*>(p.aux0 + pix * p.ldaux + n);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            dst[q] = make_float4(sigmoidf_(v[4 * q]), sigmoidf_(v[4 * q + 1]), sigmoidf_(v[4 * q + 2]), sigmoidf_(v[4 * q + 3]));
        } else {                 // r gate -> r*h as split halves
          const float4* hp = reinterpret_cast<const float4*>(p.h + pix * p.ldh + (n - Ch));
          float4 hreg[8];        // all loads first (h is read-only in this kernel): no load->store serialisation
#pragma unroll
          for (int q = 0; q < 8; ++q) hreg[q] = __ldg(hp + q);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 hv = hreg[q];
            v[4 * q + 0] = sigmoidf_(v[4 * q + 0]) * hv.x; v[4 * q + 1] = sigmoidf_(v[4 * q + 1]) * hv.y;
            v[4 * q + 2] = sigmoidf_(v[4 * q + 2]) * hv.z; v[4 * q + 3] = sigmoidf_(v[4 * q + 3]) * hv.w;
          }
          uint4* dh = reinterpret_cast<uint4*>(p.out_hi + pix * p.ldo_split + (n - Ch));
          uint4* dl = reinterpret_cast<uint4*>(p.out_lo + pix * p.ldo_split + (n - Ch));
#pragma unroll
          for (int q = 0; q < 4; ++q) split8(v + 8 * q, dh[q], dl[q]);
        }
        continue;
      }
      if (epi == RNC_EPI_GRU_Q) {
        const float4* zp = reinterpret_cast<const float4*>(p.aux0 + pix * p.ldaux + n);
        float4* hp = reinterpret_cast<float4*>(p.h + pix * p.ldh + n);
        float4 zreg[8], hreg[8];   // issue every load before the first store (hp is read-modify-write)
#pragma unroll
        for (int q = 0; q < 8; ++q) { zreg[q] = __ldg(zp + q); hreg[q] = hp[q]; }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 z = zreg[q], hv = hreg[q];
          v[4 * q + 0] = (1.f - z.x) * hv.x + z.x * tanhf(v[4 * q + 0]);
          v[4 * q + 1] = (1.f - z.y) * hv.y + z.y * tanhf(v[4 * q + 1]);
          v[4 * q + 2] = (1.f - z.z) * hv.z + z.z * tanhf(v[4 * q + 2]);
          v[4 * q + 3] = (1.f - z.w) * hv.w + z.w * tanhf(v[4 * q + 3]);
          hp[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      } else if (epi == RNC_EPI_RELU || epi == RNC_EPI_RELU_FLOW) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        if (epi == RNC_EPI_RELU_FLOW && n <= p.cout && p.cout < n + 32) {
          // append flow = coords1 - grid as channels [cout, cout+2)  (update.py:97)
          const int HW = p.H * p.W;
          const float* c1 = p.aux0 + static_cast<size_t>(b) * 2 * HW + y * p.W + x;
          const float fx = c1[0] - static_cast<float>(x), fy = c1[HW] - static_cast<float>(y);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (n + j == p.cout) v[j] = fx;
            if (n + j == p.cout + 1) v[j] = fy;
          }
        }
      } else if (epi == RNC_EPI_SIGMOID) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = sigmoidf_(v[j]);
      }
      if (p.out_f32) {
        float4* dst = reinterpret_cast<float4*>(p.out_f32 + pix * p.ldo_f32 + n);
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      if (p.out_hi) {
        uint4* dh = reinterpret_cast<uint4*>(p.out_hi + pix * p.ldo_split + n);
        uint4* dl = reinterpret_cast<uint4*>(p.out_lo + pix * p.ldo_split + n);
#pragma unroll
        for (int q = 0; q < 4; ++q) split8(v + 8 * q, dh[q], dl[q]);
      }
    }
  }

  // ------------------------------------------------------------------ teardown
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- host side
// weight plane [CoutPad][Ktot] halves: 2-D map {Ktot, CoutPad}, box {64, BN}
static bool make_w_map(CUtensorMap* m, const void* base, int ktot, int coutpad, int bn) {
  const cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)coutpad};
  const cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)bn};
  const cuuint32_t es[2] = {1, 1};
  return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN>
static int launch(const CUtensorMap* maps, const Params& p, int ntiles, int ngrid_n, cudaStream_t stream) {
  static unsigned long long done = 0;
  if (int st = ensure_dyn_smem(conv_umma_kernel<BN>, Cfg<BN>::kSmem, &done)) return st;
  conv_umma_kernel<BN><<<dim3(ntiles, ngrid_n), kThreads, Cfg<BN>::kSmem, stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
  return after_launch();
}

}  // namespace umma
}  // namespace rnc

using namespace rnc;

extern "C" int rnc_conv2d_umma_fwd(const rnc_conv_umma_desc* desc, void* stream) {
  using namespace rnc::umma;
  if (!desc) return RNC_ERR_BAD_POINTER;
  const rnc_conv_umma_desc& d = *desc;
  if (d.B <= 0 || d.H <= 0 || d.W <= 0 || d.cout <= 0 || d.c0 <= 0 || d.c1 < 0) return RNC_ERR_BAD_SHAPE;
  if (d.kh < 1 || d.kw < 1 || !(d.kh & 1) || !(d.kw & 1) || d.kh * d.kw > 49) return RNC_ERR_BAD_SHAPE;
  if ((d.ld0 & 7) || d.ld0 < d.c0 || (d.c1 > 0 && ((d.c0 % kBK) != 0 || (d.ld1 & 7) || d.ld1 < d.c1))) return RNC_ERR_BAD_SHAPE;
  if (!d.in0_hi || !d.in0_lo || (d.c1 > 0 && (!d.in1_hi || !d.in1_lo)) || !d.w_hi || !d.w_lo || !d.bias) return RNC_ERR_BAD_POINTER;
  if (!aligned16(d.in0_hi) || !aligned16(d.in0_lo) || !aligned16(d.w_hi) || !aligned16(d.w_lo) || !aligned16(d.bias)) return RNC_ERR_BAD_POINTER;
  if (d.c1 > 0 && (!aligned16(d.in1_hi) || !aligned16(d.in1_lo))) return RNC_ERR_BAD_POINTER;
  const int nblk0 = (d.c0 + kBK - 1) / kBK, nblk1 = (d.c1 + kBK - 1) / kBK, nblk = nblk0 + nblk1;
  const int ntaps = d.kh * d.kw;
  if (d.ktot != ntaps * nblk * kBK) return RNC_ERR_BAD_SHAPE;          // weight planes are [coutpad][taps * blocks * 64]
  // N tile: the whole Cout in one CTA when it fits 256 TMEM columns, else equal tiles
  int bn;
  if (d.coutpad <= 32) bn = 32; else if (d.coutpad <= 64) bn = 64; else if (d.coutpad <= 128) bn = 128;
  else if (d.coutpad == 192 || d.coutpad % 192 == 0) bn = 192; else bn = 256;
  if (d.coutpad % bn != 0 || d.coutpad < d.cout) return RNC_ERR_BAD_SHAPE;
  if (d.epilogue == RNC_EPI_RELU_FLOW && (d.coutpad < d.cout + 2 || !d.aux0 || !d.out_hi)) return RNC_ERR_BAD_SHAPE;
  if (d.out_hi && (!d.out_lo || (d.ldo_split & 7) || !aligned16(d.out_hi) || !aligned16(d.out_lo))) return RNC_ERR_BAD_POINTER;
  if (d.out_f32 && ((d.ldo_f32 & 3) || !aligned16(d.out_f32))) return RNC_ERR_BAD_POINTER;
  switch (d.epilogue) {
    case RNC_EPI_LINEAR: case RNC_EPI_RELU: case RNC_EPI_SIGMOID: case RNC_EPI_RELU_FLOW:
      if (!d.out_f32 && !d.out_hi) return RNC_ERR_BAD_POINTER;
      if ((d.cout % 32) != 0 && d.epilogue != RNC_EPI_RELU_FLOW) {
        // the epilogue stores whole 32-channel chunks: the destination must have room for the padded tail
        const int cpad = (d.cout + 31) / 32 * 32;
        if ((d.out_f32 && d.ldo_f32 < cpad) || (d.out_hi && d.ldo_split < cpad)) return RNC_ERR_BAD_SHAPE;
      }
      break;
    case RNC_EPI_GRU_ZR:
      if (!d.out_hi || !d.aux0 || !d.h || (d.cout % 64) != 0 || (d.ldaux & 3) || (d.ldh & 3)) return RNC_ERR_BAD_POINTER;
      break;
    case RNC_EPI_GRU_Q:
      if (!d.aux0 || !d.h || (d.cout % 32) != 0 || (d.ldaux & 3) || (d.ldh & 3)) return RNC_ERR_BAD_POINTER;
      break;
    default: return RNC_ERR_UNSUPPORTED;
  }
  if (!encode_fn()) return RNC_ERR_UNSUPPORTED;

  // pixel tile: TW x TH = 128 with TW the smallest power of two covering min(W, 128)
  int TW = 8;
  while (TW < d.W && TW < kBM) TW <<= 1;
  const int TH = kBM / TW;
  Params p;
  p.B = d.B; p.H = d.H; p.W = d.W; p.TW = TW; p.TH = TH;
  p.tiles_x = (d.W + TW - 1) / TW; p.tiles_y = (d.H + TH - 1) / TH;
  p.kw = d.kw; p.ph = d.kh / 2; p.pw = d.kw / 2; p.ntaps = ntaps;
  p.nblk0 = nblk0; p.nblk = nblk;
  p.cout = d.cout; p.epilogue = d.epilogue; p.unscale = d.unscale; p.bias = d.bias;
  p.out_f32 = d.out_f32; p.ldo_f32 = d.ldo_f32;
  p.out_hi = static_cast<__half*>(d.out_hi); p.out_lo = static_cast<__half*>(d.out_lo); p.ldo_split = d.ldo_split;
  p.h = d.h; p.ldh = d.ldh; p.aux0 = d.aux0; p.ldaux = d.ldaux;

  CUtensorMap maps[6];
  bool ok = make_act_map(&maps[0], d.in0_hi, d.c0, d.ld0, d.B, d.H, d.W, TW, TH) &&
            make_act_map(&maps[1], d.in0_lo, d.c0, d.ld0, d.B, d.H, d.W, TW, TH);
  if (d.c1 > 0) {
    ok = ok && make_act_map(&maps[2], d.in1_hi, d.c1, d.ld1, d.B, d.H, d.W, TW, TH) &&
         make_act_map(&maps[3], d.in1_lo, d.c1, d.ld1, d.B, d.H, d.W, TW, TH);
  } else {
    maps[2] = maps[0]; maps[3] = maps[1];
  }
  ok = ok && make_w_map(&maps[4], d.w_hi, d.ktot, d.coutpad, bn) && make_w_map(&maps[5], d.w_lo, d.ktot, d.coutpad, bn);
  if (!ok) return RNC_ERR_BAD_SHAPE;

  const int ntiles = d.B * p.tiles_x * p.tiles_y, gn = d.coutpad / bn;
  cudaStream_t s = as_stream(stream);
  switch (bn) {
    case 32: return launch<32>(maps, p, ntiles, gn, s);
    case 64: return launch<64>(maps, p, ntiles, gn, s);
    case 128: return launch<128>(maps, p, ntiles, gn, s);
    case 192: return launch<192>(maps, p, ntiles, gn, s);
    default: return launch<256>(maps, p, ntiles, gn, s);
  }
}
