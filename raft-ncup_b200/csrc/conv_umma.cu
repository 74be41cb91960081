// tcgen05 implicit-GEMM convolution (v3): persistent CTAs, TMA-staged 128B-swizzled operand rings, tcgen05.mma
// (kind::f16, M128 x N<=256 x K16) issued by one thread, fp32 accumulators double-buffered in TMEM so that the epilogue
// of tile i overlaps the main loop of tile i+1.
// Replaces: core/update.py:33-60 (SepConvGRU), :79-97 (BasicMotionEncoder), :6-14 (FlowHead), :123-126 (mask head), the 3x3
// layers of core/interp_weights_est.py:10-47 and (stride 1/2, residual epilogue) the convolutions of
// core/extractor.py:6-56,118-192.
//
// fp32-faithful on fp16 tensor cores: every activation x and weight w is an exact-sum pair of halves (x = x_hi + x_lo;
// weights pre-scaled by a power of two so w_lo stays normal) and each K step issues x_hi*w_hi + x_hi*w_lo + x_lo*w_hi
// (3 MMAs): plain TF32/bf16 operands miss the 1e-3 EPE bar by 17-60x.
//
// GEMM view: M = 128 output pixels, N = Cout tile, K = taps x 64-channel blocks.  Zero padding is free: A tiles are
// 4-D TMA boxes whose out-of-image elements are zero-filled.  Three A-staging modes cut the L2->SM traffic of the taps:
//   TAP      one 128-row box per filter tap (any geometry, strides 1 and 2 via TMA element strides)
//   ROWHALO  tile = 1 image row x 128 px: one (128 + 8)-row box per (ky, channel block); the kw horizontal taps are
//            operand descriptors shifted by kx rows (descriptor base_offset carries the swizzle phase)
//   COLHALO  kw == 1: tile = 8 x 16 px with a vertical halo, the kh taps are descriptors shifted by ky*16 rows
// A fourth input form, RNC_CONV_WINDOW, reads in0 through a sliding-window tensor map (dimension 1 steps fewer bytes than
// dimension 0 spans): the TMA unit builds im2col rows of a small-Cin layer (the encoders' 7x7/2 stem) without a copy.
// Epilogue: TMEM -> registers -> per-warp swizzled staging in shared memory -> cp.async.bulk.tensor stores (split halves,
// fp32 outputs, the GRU state); fused InstanceNorm sums read the staged chunk column-wise.
#include "umma_ptx.cuh"

namespace rnc {
namespace umma {

constexpr int kThreads = 320;        // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue (2 per lane group)
constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kMaxSA = 4, kMaxSB = 12;
constexpr int kStageWarp = 4096;     // epilogue staging per epilogue warp: 32 px x 32 ch as [hi 2 KB | lo 2 KB] halves or 4 KB of fp32
constexpr int kStageBytes = 8 * kStageWarp;
constexpr int kRingBudget = 184 * 1024;   // A + B rings; + barriers, statistics slots and the epilogue staging = 227 KB
enum { MODE_TAP = 0, MODE_ROWHALO = 1, MODE_COLHALO = 2 };

struct Params {
  int B, H, W;                        // output geometry
  int TW, TH, tiles_x, tiles_y, ntiles, ntn;
  int mode, kh, kw, ph, pw, stride;   // stride: y (and x unless sx overrides)
  int sx;                             // x stride (1 for a window view, whose positions already step by the stride)
  int nblk0, nblk;                    // K blocks (128 bytes of channels: 64 halves or 32 TF32 words) in segment 0 / total
  int bk;                             // channels per K block: 64 (fp16 hi/lo operands) or 32 (TF32 hi/lo operands)
  int tf32;                           // operands are fp32 planes consumed as TF32 (kind::tf32): training-path layers
  int a_plane, SA, SB;                // bytes per A half-plane stage (rows*128, 1024-aligned), ring depths
  int resident_b;                     // the layer's whole weight matrix fits the B ring: loaded once per CTA, never released
  int probe_nob;                      // developer probe (RNC_CONV_PROBE_NOB=1): skip the weight loads after the first ring fill
  int cout, epilogue;
  float unscale;
  const float* bias;
  float* out_f32; int ldo_f32;
  __half* out_hi; __half* out_lo; int ldo_split;
  float* h; int ldh;
  float* aux0; int ldaux;
  const float* res; int ldres;
  const float* add; int ldadd;        // optional fp32 addend of the pre-activation, [pixel][ldadd]
  int aux_blocked, out_blocked;       // aux0 + add / out_f32 in the tile-blocked layout [tile][channel][128 px] (coalesced for thread = pixel)
  double* stats;                      // [B][cout][2]: per-(image, channel) sum / sum of squares of the outputs, accumulated
};

// exact hi/lo split of 8 floats into two 16-byte vectors of halves
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pair(v[2 * i], v[2 * i + 1], h[i], l[i]);
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int BN>
struct Cfg {
  static constexpr int kBTile = BN * kBK * 2;                     // bytes per half-plane of weights
  // Each tile owns TWO accumulators: `main` takes the x_hi*w_hi products, `corr` the two 2^-11-smaller cross terms.
  // The tensor core truncates on every accumulate, so the error grows with the number of full-magnitude adds; keeping
  // the cross terms out of `main` cuts it 3x (measured: 4.8e-5 -> 1.7e-5 on a 3x3x256 layer).  BN <= 128 double-buffers
  // the pair (epilogue overlaps the next tile's main loop); wider tiles fit only one pair in the 512 TMEM columns.
  static constexpr int kBufs = BN <= 128 ? 2 : 1;
  static constexpr int kTmemCols = BN <= 32 ? 128 : BN <= 64 ? 256 : 512;
};

// PAIR: two CTAs of a cluster (adjacent pixel tiles, same weights) run each MMA together (cta_group::2, M = 256): every CTA
// stages its own A tile and HALF of the weight tile, so the shared-memory fill and operand-read traffic per SM drop by the
// weight share — the resource that bounds the single-CTA form.  The leader (rank 0) issues all MMAs.
// EC: epilogue class, compiled separately so that a launch carries only its own epilogue code (the union of all epilogues cost
// the gate layers instruction-cache misses — 18 % of the epilogue warps' stall samples — and registers): 0 = plain layers
// (linear / relu / sigmoid), 1 = GRU z|r gates, 2 = GRU q, 3 = the rest (flow append, residual tail, tanh|relu head,
// coords update, fused InstanceNorm sums).
enum { EC_PLAIN = 0, EC_GRU_ZR = 1, EC_GRU_Q = 2, EC_MISC = 3 };

template <int BN, bool PAIR, int EC>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap mA0h, const __grid_constant__ CUtensorMap mA0l,
                 const __grid_constant__ CUtensorMap mA1h, const __grid_constant__ CUtensorMap mA1l,
                 const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl,
                 const __grid_constant__ CUtensorMap mOh, const __grid_constant__ CUtensorMap mOl,
                 const __grid_constant__ CUtensorMap mOf, const Params p) {
  using C = Cfg<BN>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_stage = 2 * p.a_plane;
  unsigned char* sA = smem;
  unsigned char* sB = smem + p.SA * a_stage;
  // Pair form of narrow tiles (BN <= 64): an MMA of 64 columns is bound by the issue rate, so the two products that share x_hi
  // become ONE instruction of 2*BN columns — the leader stages all BN rows of w_hi, its peer all BN rows of w_lo at the same
  // offset X (a cta_group::2 MMA takes half of its N columns from each CTA: columns [0, BN) = x_hi*w_hi -> main, [BN, 2BN) =
  // x_hi*w_lo -> corr, adjacent in TMEM) — and x_lo*w_hi reads each CTA's half of w_hi from a second region Y.
  constexpr bool kFusedPair = PAIR && BN <= 64;
  constexpr int kBStageBytes = kFusedPair ? C::kBTile + C::kBTile / 2 : PAIR ? C::kBTile : 2 * C::kBTile;   // weight bytes per K block and CTA
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + p.SB * kBStageBytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kMaxSA;
  uint64_t* b_full = a_empty + kMaxSA;
  uint64_t* b_empty = b_full + kMaxSB;
  uint64_t* acc_full = b_empty + kMaxSB;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  double* stat_acc = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(bars) + 512);   // [8 warps][kHalfN][32 lanes][2], BN <= 128 only
  // epilogue staging (TMA-store source), 1024-aligned so the 64B / 128B swizzle patterns are functions of the buffer offset
  unsigned char* stage_base = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(bars) + 512 + (BN <= 128 ? 8192 : 0) + 1023) & ~uintptr_t(1023));

  pdl_trigger();                       // the next kernel in the stream may begin its prologue on SMs this grid has left
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tpi = p.tiles_x * p.tiles_y;
  const int G = p.mode == MODE_TAP ? p.kh * p.kw : p.mode == MODE_ROWHALO ? p.kh : 1;     // A-stage groups per channel block
  const int T = p.mode == MODE_TAP ? 1 : p.mode == MODE_ROWHALO ? p.kw : p.kh;            // taps sharing one A stage
  // work items: (pixel tile [pair], column tile); a pair's two CTAs take adjacent pixel tiles of the same item
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int ptiles = PAIR ? (p.ntiles + 1) / 2 : p.ntiles;
  const int items = ptiles * p.ntn;
  const int item0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int item_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  constexpr int kBHalf = PAIR ? C::kBTile / 2 : C::kBTile;          // bytes of one weight half-plane staged by this CTA

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], PAIR ? 16 : 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) { if (PAIR) tmem_alloc_pair(tmem_slot, C::kTmemCols); else tmem_alloc(tmem_slot, C::kTmemCols); }
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();        // the peer's barriers exist before anything signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                          // everything above overlapped the predecessor's tail; its results are visible from here

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp runs the loop so that addresses / coordinates stay in uniform registers (UTMALDG and UTCHMMA take
    // uniform operands; a loop under `if (lane == 0)` forces an R2UR chain in front of every issue); one elected lane issues.
    {
      // pair: both CTAs' loads complete on the LEADER's full barriers (it alone waits for operands); each CTA recycles its
      // own stages when the leader's MMA completion reaches its own empty barriers
      const uint32_t a_full_l = PAIR ? mapa_u32(smem_u32(a_full), 0) : 0u, b_full_l = PAIR ? mapa_u32(smem_u32(b_full), 0) : 0u;
      // ring positions as (slot, parity) counters: a runtime modulo per stage costs ~100 cycles of dependent integer code in
      // these single-thread loops, which the 32-cycle MMAs of narrow tiles do not hide
      int sa_n = 0, pa_n = 0, sb_n = 0, pb_n = 0, b_it = 0;
      for (int item = item0; item < items; item += item_step) {
        const int ptile = item / p.ntn, n0 = (item - ptile * p.ntn) * BN;
        const int tile = PAIR ? ptile * 2 + static_cast<int>(rank) : ptile;     // a ghost tile (odd count) loads zero-filled boxes
        const int b = tile / tpi, tr = tile - b * tpi;
        const int y0 = (tr / p.tiles_x) * p.TH, x0 = (tr % p.tiles_x) * p.TW;
        for (int g = 0; g < G; ++g) {
          int cx, cy;
          if (p.mode == MODE_TAP) { cx = x0 * p.sx + g % p.kw - p.pw; cy = y0 * p.stride + g / p.kw - p.ph; }
          else if (p.mode == MODE_ROWHALO) { cx = x0 - p.pw; cy = y0 + g - p.ph; }
          else { cx = x0; cy = y0 - p.ph; }
          for (int cb = 0; cb < p.nblk; ++cb) {
            const int sa = sa_n, pa = pa_n;
            if (++sa_n == p.SA) { sa_n = 0; pa_n ^= 1; }
            mbar_wait(&a_empty[sa], pa ^ 1);
            const bool seg0 = cb < p.nblk0;
            const int c = (seg0 ? cb : cb - p.nblk0) * p.bk;
            if (elect_one()) {
              if (PAIR) {
                if (leader) mbar_expect_tx(&a_full[sa], 2 * a_stage);
                tma_load_4d_pair(sA + sa * a_stage, seg0 ? &mA0h : &mA1h, a_full_l + sa * 8, c, cx, cy, b);
                tma_load_4d_pair(sA + sa * a_stage + p.a_plane, seg0 ? &mA0l : &mA1l, a_full_l + sa * 8, c, cx, cy, b);
              } else {
                mbar_expect_tx(&a_full[sa], a_stage);
                tma_load_4d(sA + sa * a_stage, seg0 ? &mA0h : &mA1h, &a_full[sa], c, cx, cy, b);
                tma_load_4d(sA + sa * a_stage + p.a_plane, seg0 ? &mA0l : &mA1l, &a_full[sa], c, cx, cy, b);
              }
            }
            __syncwarp();
            for (int t = 0; t < T; ++t) {
              const int tap = p.mode == MODE_TAP ? g : p.mode == MODE_ROWHALO ? g * p.kw + t : t;
              const int sb = sb_n, pb = pb_n;
              if (++sb_n == p.SB) { sb_n = 0; pb_n ^= 1; }
              ++b_it;
              if (p.resident_b && item != item0) continue;     // weights already resident
              mbar_wait(&b_empty[sb], pb ^ 1);
              const int kcol = (tap * p.nblk + cb) * p.bk;
              if (elect_one()) {
                if (kFusedPair) {            // X: all BN rows of w_hi (leader) / w_lo (peer) as two half boxes; Y: own half of w_hi
                  const CUtensorMap* mx = leader ? &mBh : &mBl;
                  unsigned char* st = sB + sb * kBStageBytes;
                  if (leader) mbar_expect_tx(&b_full[sb], 2 * kBStageBytes);
                  tma_load_2d_pair(st, mx, b_full_l + sb * 8, kcol, n0);
                  tma_load_2d_pair(st + C::kBTile / 2, mx, b_full_l + sb * 8, kcol, n0 + BN / 2);
                  tma_load_2d_pair(st + C::kBTile, &mBh, b_full_l + sb * 8, kcol, n0 + static_cast<int>(rank) * (BN / 2));
                } else if (PAIR) {           // this CTA's half of the weight rows, hi and lo planes
                  const int nrow = n0 + static_cast<int>(rank) * (BN / 2);
                  if (leader) mbar_expect_tx(&b_full[sb], 2 * kBStageBytes);
                  tma_load_2d_pair(sB + sb * kBStageBytes, &mBh, b_full_l + sb * 8, kcol, nrow);
                  tma_load_2d_pair(sB + sb * kBStageBytes + kBHalf, &mBl, b_full_l + sb * 8, kcol, nrow);
                } else if (p.probe_nob && b_it > p.SB) {
                  mbar_arrive(&b_full[sb]);
                } else {
                  mbar_expect_tx(&b_full[sb], 2 * C::kBTile);
                  tma_load_2d(sB + sb * 2 * C::kBTile, &mBh, &b_full[sb], kcol, n0);
                  tma_load_2d(sB + sb * 2 * C::kBTile + C::kBTile, &mBl, &b_full[sb], kcol, n0);
                }
              }
              __syncwarp();
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane issues)
    if (leader) {
      // instruction descriptor: D = F32 (bit 4), A/B format bits [7,10) / [10,13): 0 = F16, 2 = TF32; N >> 3 at 17, M >> 4 at 24
      const uint32_t fmt = p.tf32 ? ((2u << 7) | (2u << 10)) : 0u;
      const uint32_t idesc = (1u << 4) | fmt | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>((PAIR ? 2 * kBM : kBM) >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | fmt | (static_cast<uint32_t>((2 * BN <= 256 ? 2 * BN : BN) >> 3) << 17) | (static_cast<uint32_t>(kBM >> 4) << 24);
      const bool tf32 = p.tf32 != 0;
      // Operand descriptors differ only in their address field (bits [0,14) = byte address >> 4, stages and tap shifts are
      // multiples of 16 bytes and shared memory ends below 256 KB): stage 0's descriptor + one multiply-add per use, instead
      // of rebuilding the 64-bit word from the address in this single-thread loop.  A descriptor whose start is shifted by kx
      // rows needs no base_offset: with 1024-byte-aligned stages the swizzle is a function of the absolute address.
      const uint64_t a_desc0 = smem_desc_sw128(smem_u32(sA)), b_desc0 = smem_desc_sw128(smem_u32(sB));
      const uint32_t a_step = static_cast<uint32_t>(a_stage) >> 4, a_lo = static_cast<uint32_t>(p.a_plane) >> 4;
      const uint32_t t_step = static_cast<uint32_t>(p.mode == MODE_ROWHALO ? 1 : p.mode == MODE_COLHALO ? p.TW : 0) * 8u;   // tap shift: rows * 128 B >> 4
      int sa_n = 0, pa_n = 0, sb_n = 0, pb_n = 0, t_it = 0;
      for (int item = item0; item < items; item += item_step, ++t_it) {
        const int buf = t_it % C::kBufs, use = t_it / C::kBufs;
        mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t d_main = tmem_base + buf * 2 * BN, d_corr = d_main + BN;
        uint32_t acc = 0;
        for (int g = 0; g < G; ++g)
          for (int cb = 0; cb < p.nblk; ++cb) {
            const int sa = sa_n, pa = pa_n;
            if (++sa_n == p.SA) { sa_n = 0; pa_n ^= 1; }
            mbar_wait(&a_full[sa], pa);
            for (int t = 0; t < T; ++t) {
              const int sb = sb_n, pb = pb_n;
              if (++sb_n == p.SB) { sb_n = 0; pb_n ^= 1; }
              mbar_wait(&b_full[sb], p.resident_b ? 0 : pb);
              tcgen05_fence_after();
              const uint64_t ah = a_desc0 + (sa * a_step + t * t_step), al = ah + a_lo;
              const uint64_t bh = b_desc0 + sb * (kBStageBytes >> 4), bl = bh + (kBHalf >> 4);
              if (elect_one()) {
                if (kFusedPair) {
                  const uint64_t bx = bh, by = bh + (C::kBTile >> 4);
                  const uint32_t idesc_x = (1u << 4) | fmt | (static_cast<uint32_t>((2 * BN) >> 3) << 17) | (static_cast<uint32_t>((2 * kBM) >> 4) << 24);
#pragma unroll
                  for (int k = 0; k < kBK / 16; ++k) {
                    if (tf32) {
                      umma_tf32_pair(d_main, ah + 2 * k, bx + 2 * k, idesc_x, k == 0 ? acc : 1u);     // main | corr
                      umma_tf32_pair(d_corr, al + 2 * k, by + 2 * k, idesc, 1);
                    } else {
                      umma_f16_pair(d_main, ah + 2 * k, bx + 2 * k, idesc_x, k == 0 ? acc : 1u);      // main | corr
                      umma_f16_pair(d_corr, al + 2 * k, by + 2 * k, idesc, 1);
                    }
                  }
                  if (!p.resident_b) umma_commit_pair(&b_empty[sb]);
                } else if (PAIR) {
                  // M = 256 across the pair: each CTA's tensor core reads its own A rows and both CTAs' weight halves
#pragma unroll
                  for (int k = 0; k < kBK / 16; ++k) {     // 4 K steps of 32 bytes per 128-byte block (16 halves or 8 TF32 words)
                    if (tf32) {
                      umma_tf32_pair(d_main, ah + 2 * k, bh + 2 * k, idesc, k == 0 ? acc : 1u);
                      umma_tf32_pair(d_corr, ah + 2 * k, bl + 2 * k, idesc, k == 0 ? acc : 1u);
                      umma_tf32_pair(d_corr, al + 2 * k, bh + 2 * k, idesc, 1);
                    } else {
                      umma_f16_pair(d_main, ah + 2 * k, bh + 2 * k, idesc, k == 0 ? acc : 1u);
                      umma_f16_pair(d_corr, ah + 2 * k, bl + 2 * k, idesc, k == 0 ? acc : 1u);
                      umma_f16_pair(d_corr, al + 2 * k, bh + 2 * k, idesc, 1);
                    }
                  }
                  if (!p.resident_b) umma_commit_pair(&b_empty[sb]);
                } else {
                  if (BN <= 128) {
                    // The hi and lo weight planes are adjacent in the stage and `main`, `corr` adjacent in TMEM, so
                    // x_hi * [w_hi | w_lo] is ONE instruction of 2*BN columns: x_hi is fetched from shared memory once
                    // instead of twice (a 128-column instruction needs the full 128 B/clk of shared-memory bandwidth).
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                      if (tf32) {
                        umma_tf32(d_main, ah + 2 * k, bh + 2 * k, idesc2, k == 0 ? acc : 1u);
                        umma_tf32(d_corr, al + 2 * k, bh + 2 * k, idesc, 1);
                      } else {
                        umma_f16(d_main, ah + 2 * k, bh + 2 * k, idesc2, k == 0 ? acc : 1u);
                        umma_f16(d_corr, al + 2 * k, bh + 2 * k, idesc, 1);
                      }
                    }
                  } else {
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                      if (tf32) {
                        umma_tf32(d_main, ah + 2 * k, bh + 2 * k, idesc, k == 0 ? acc : 1u);
                        umma_tf32(d_corr, ah + 2 * k, bl + 2 * k, idesc, k == 0 ? acc : 1u);
                        umma_tf32(d_corr, al + 2 * k, bh + 2 * k, idesc, 1);
                      } else {
                        umma_f16(d_main, ah + 2 * k, bh + 2 * k, idesc, k == 0 ? acc : 1u);
                        umma_f16(d_corr, ah + 2 * k, bl + 2 * k, idesc, k == 0 ? acc : 1u);
                        umma_f16(d_corr, al + 2 * k, bh + 2 * k, idesc, 1);
                      }
                    }
                  }
                  if (!p.resident_b) umma_commit(&b_empty[sb]);
                }
              }
              __syncwarp();
              acc = 1;
            }
            if (elect_one()) { if (PAIR) umma_commit_pair(&a_empty[sa]); else umma_commit(&a_empty[sa]); }
            __syncwarp();
          }
        if (elect_one()) { if (PAIR) umma_commit_pair(&acc_full[buf]); else umma_commit(&acc_full[buf]); }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> registers -> global
    const int lg = warp & 3;                       // TMEM lane group this warp may access
    const int ml = lg * 32 + lane;                 // row of the tile = pixel
    const int epi = p.epilogue;
    constexpr int kChunksN = BN / 32, kHalfN = (kChunksN + 1) / 2;
    const int cc0 = (warp - 2) < 4 ? 0 : kHalfN, cc1 = (warp - 2) < 4 ? kHalfN : kChunksN;
    const bool use_stats = EC == EC_MISC && BN <= 128 && p.stats != nullptr;
    double* my_acc = stat_acc + (static_cast<size_t>(warp - 2) * kHalfN * 32 + lane) * 2;   // [chunk] stride 64 doubles
    int acc_b = -1, acc_n0 = 0;
    auto flush_stats = [&]() {
      if (acc_b < 0) return;
      for (int ci = 0; ci < cc1 - cc0; ++ci) {
        const int ch = acc_n0 + (cc0 + ci) * 32 + lane;
        if (ch < p.cout) {
          double* dst = p.stats + (static_cast<size_t>(acc_b) * p.cout + ch) * 2;
          atomicAdd(dst, my_acc[ci * 64]);
          atomicAdd(dst + 1, my_acc[ci * 64 + 1]);
        }
        my_acc[ci * 64] = 0.0; my_acc[ci * 64 + 1] = 0.0;
      }
    };
    if (use_stats) {
      for (int ci = 0; ci < kHalfN; ++ci) { my_acc[ci * 64] = 0.0; my_acc[ci * 64 + 1] = 0.0; }
    }
    const uint32_t acc_empty_l = PAIR ? mapa_u32(smem_u32(acc_empty), 0) : 0u;     // the leader's MMA warp waits for both CTAs
    // Outputs leave through shared memory and TMA stores: a warp's 32 px x 32 ch chunk is one box of the output tensor
    // (full 64- / 128-byte rows per pixel instead of 16-byte pieces per thread; image borders and ghost tiles are clipped by
    // the TMA unit).  The staging rows are written with the map's swizzle (64B for halves, 128B for fp32): conflict-free.
    unsigned char* stg = stage_base + (warp - 2) * kStageWarp;
    bool stg_busy = false;                         // warp-uniform: a TMA store may still be reading the staging buffer
    const int lgx = (lg * 32) % p.TW, lgy = (lg * 32) / p.TW;
    int bx = 0, by = 0, bb = 0;                    // box origin of this warp's lane group in the current tile
    auto stage_free = [&]() {
      if (stg_busy) { if (lane == 0) bulk_wait_read0(); __syncwarp(); stg_busy = false; }
    };
    auto store_split = [&](const float* v, int ch) {           // hi/lo halves of v[0..32) -> channels [ch, ch+32) of out_hi / out_lo
      stage_free();
      uint4* sh = reinterpret_cast<uint4*>(stg + lane * 64);
      uint4* sl = reinterpret_cast<uint4*>(stg + 2048 + lane * 64);
      const int sw = (lane >> 1) & 3;
#pragma unroll
      for (int q = 0; q < 4; ++q) split8(v + 8 * q, sh[q ^ sw], sl[q ^ sw]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&mOh, stg, ch, bx, by, bb);
        tma_store_4d(&mOl, stg + 2048, ch, bx, by, bb);
        bulk_commit();
      }
      stg_busy = true;
    };
    auto store_f32 = [&](const float* v, int ch) {             // fp32 v[0..32) -> channels [ch, ch+32) of h (q gate) / out_f32
      stage_free();
      float4* sf = reinterpret_cast<float4*>(stg + lane * 128);
      const int sw = lane & 7;
#pragma unroll
      for (int q = 0; q < 8; ++q) sf[q ^ sw] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { tma_store_4d(&mOf, stg, ch, bx, by, bb); bulk_commit(); }
      stg_busy = true;
    };
    int t_it = 0;
    for (int item = item0; item < items; item += item_step, ++t_it) {
      const int ptile = item / p.ntn, n0 = (item - ptile * p.ntn) * BN;
      const int tile = PAIR ? ptile * 2 + static_cast<int>(rank) : ptile;
      const int b = tile / tpi, tr = tile - b * tpi;
      const bool ghost = tile >= p.ntiles;                     // odd tile count: the pair's second CTA idles through this item
      if (use_stats && !ghost && (b != acc_b || n0 != acc_n0)) { flush_stats(); acc_b = b; acc_n0 = n0; }
      const int ty0 = (tr / p.tiles_x) * p.TH, tx0 = (tr % p.tiles_x) * p.TW;
      const int y = ty0 + ml / p.TW, x = tx0 + ml % p.TW;
      const bool valid = !ghost && y < p.H && x < p.W;
      bx = tx0 + lgx; by = ty0 + lgy; bb = b;                  // a ghost tile has b >= B: its boxes are clipped whole
      // tile-blocked tensors (p.aux_blocked / p.out_blocked): element (tile, channel c, row ml) at ((tile * ld + c) * 128 + ml),
      // so a warp's 32 pixels are contiguous per channel
      const size_t pix = (static_cast<size_t>(b) * p.H + y) * p.W + x;
      const int buf = t_it % C::kBufs, use = t_it / C::kBufs;
      mbar_wait(&acc_full[buf], use & 1);
      tcgen05_fence_after();
      // two epilogue warps per TMEM lane group: the first takes the lower half of the 32-channel chunks, the second the rest
#pragma unroll 1
      for (int cc = cc0; cc < cc1; ++cc) {
        const int n = n0 + cc * 32;
        if (n >= p.cout + (EC == EC_MISC && epi == RNC_EPI_RELU_FLOW ? 2 : 0)) break;   // warp-uniform
        uint32_t r[32], rc[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + buf * 2 * BN + cc * 32;
        tmem_ld32(taddr, r);
        tmem_ld32(taddr + BN, rc);
        if (ghost) continue;                                           // warp-uniform: nothing to add to the statistics either
        float v[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n) + q);
          v[4 * q + 0] = fmaf(__uint_as_float(r[4 * q + 0]) + __uint_as_float(rc[4 * q + 0]), p.unscale, bv.x);
          v[4 * q + 1] = fmaf(__uint_as_float(r[4 * q + 1]) + __uint_as_float(rc[4 * q + 1]), p.unscale, bv.y);
          v[4 * q + 2] = fmaf(__uint_as_float(r[4 * q + 2]) + __uint_as_float(rc[4 * q + 2]), p.unscale, bv.z);
          v[4 * q + 3] = fmaf(__uint_as_float(r[4 * q + 3]) + __uint_as_float(rc[4 * q + 3]), p.unscale, bv.w);
        }
        if (p.add != nullptr && valid) {
          // hoisted part of the layer (input channels that are constant across calls), computed once by another launch
          if (p.aux_blocked) {
            const float* ap = p.add + (static_cast<size_t>(tile) * p.ldadd + n) * kBM + ml;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __ldg(ap + static_cast<size_t>(j) * kBM);
          } else {
            const float4* ap = reinterpret_cast<const float4*>(p.add + pix * p.ldadd + n);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 a = __ldg(ap + q);
              v[4 * q] += a.x; v[4 * q + 1] += a.y; v[4 * q + 2] += a.z; v[4 * q + 3] += a.w;
            }
          }
        }
        if (use_stats) {
          // InstanceNorm statistics of this layer's output (extractor.py:128-129), fused here instead of a pass over the fp32
          // tensor.  The chunk is staged in shared memory for its TMA store anyway ([pixel][32 ch], 128B-swizzled), which is
          // the transposition the reduction over pixels needs: lane = channel walks its column (conflict-free: a row's 32
          // words are a permutation of the 32 banks), fp32 partial sums over the warp's 32 pixels, accumulated in fp64 per
          // lane in shared memory across the CTA's tiles and flushed with fp64 atomics when the image changes.  Pixels beyond
          // the image are staged as zeros (the TMA store clips them).
          if (!valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          store_f32(v, n);                       // stats layers are RNC_EPI_LINEAR with a channel-last fp32 output (host check)
          float s1 = 0.f, s2 = 0.f;
          const int cw = lane >> 2, ce = lane & 3;
#pragma unroll
          for (int px = 0; px < 32; ++px) {
            const float x = *reinterpret_cast<const float*>(stg + px * 128 + ((cw ^ (px & 7)) << 4) + ce * 4);
            s1 += x;
            s2 = fmaf(x, x, s2);
          }
          my_acc[(cc - cc0) * 64] += static_cast<double>(s1);
          my_acc[(cc - cc0) * 64 + 1] += static_cast<double>(s2);
          continue;
        }
        if (EC == EC_GRU_ZR) {
          const int Ch = p.cout >> 1;
          if (n < Ch) {            // z gate -> fp32 aux buffer
            if (valid) {
              if (p.aux_blocked) {
                float* dst = p.aux0 + (static_cast<size_t>(tile) * p.ldaux + n) * kBM + ml;
#pragma unroll
                for (int j = 0; j < 32; ++j) dst[static_cast<size_t>(j) * kBM] = sigmoid_fast(v[j]);
              } else {
                float4* dst = reinterpret_cast<float4*>(p.aux0 + pix * p.ldaux + n);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  dst[q] = make_float4(sigmoid_fast(v[4 * q]), sigmoid_fast(v[4 * q + 1]), sigmoid_fast(v[4 * q + 2]), sigmoid_fast(v[4 * q + 3]));
              }
            }
          } else {                 // r gate -> r*h as split halves
            const float4* hp = reinterpret_cast<const float4*>(p.h + pix * p.ldh + (n - Ch));
            float4 hreg[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) hreg[q] = valid ? __ldg(hp + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 hv = hreg[q];
              v[4 * q + 0] = sigmoid_fast(v[4 * q + 0]) * hv.x; v[4 * q + 1] = sigmoid_fast(v[4 * q + 1]) * hv.y;
              v[4 * q + 2] = sigmoid_fast(v[4 * q + 2]) * hv.z; v[4 * q + 3] = sigmoid_fast(v[4 * q + 3]) * hv.w;
            }
            store_split(v, n - Ch);
          }
          continue;
        }
        if (EC == EC_MISC && epi == RNC_EPI_FLOW_DELTA) {
          // FlowHead.conv2 + `coords1 = coords1 + delta_flow` (update.py:14, raft_nc_dbl.py:157); only channels 0,1 are real
          if (valid) {
            const int HW = p.H * p.W;
            const size_t i0 = static_cast<size_t>(b) * 2 * HW + y * p.W + x;
            p.aux0[i0] += v[0];
            p.aux0[i0 + HW] += v[1];
            if (p.out_f32) { p.out_f32[i0] = v[0]; p.out_f32[i0 + HW] = v[1]; }
          }
          continue;
        }
        bool want_f32 = p.out_f32 != nullptr;
        if (EC == EC_GRU_Q) {
          const float4* hp = reinterpret_cast<const float4*>(p.h + pix * p.ldh + n);
          float4 zreg[8], hreg[8];
          if (!valid) {
#pragma unroll
            for (int q = 0; q < 8; ++q) { zreg[q] = make_float4(0.f, 0.f, 0.f, 0.f); hreg[q] = zreg[q]; }
          } else {
            if (p.aux_blocked) {
              const float* zb = p.aux0 + (static_cast<size_t>(tile) * p.ldaux + n) * kBM + ml;
#pragma unroll
              for (int q = 0; q < 8; ++q)
                zreg[q] = make_float4(__ldg(zb + static_cast<size_t>(4 * q) * kBM), __ldg(zb + static_cast<size_t>(4 * q + 1) * kBM),
                                      __ldg(zb + static_cast<size_t>(4 * q + 2) * kBM), __ldg(zb + static_cast<size_t>(4 * q + 3) * kBM));
            } else {
              const float4* zp = reinterpret_cast<const float4*>(p.aux0 + pix * p.ldaux + n);
#pragma unroll
              for (int q = 0; q < 8; ++q) zreg[q] = __ldg(zp + q);
            }
            // plain loads: h is rewritten by this kernel (each chunk is read before its own TMA store is issued)
#pragma unroll
            for (int q = 0; q < 8; ++q) hreg[q] = hp[q];
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 z = zreg[q], hv = hreg[q];
            v[4 * q + 0] = (1.f - z.x) * hv.x + z.x * tanh_fast(v[4 * q + 0]);
            v[4 * q + 1] = (1.f - z.y) * hv.y + z.y * tanh_fast(v[4 * q + 1]);
            v[4 * q + 2] = (1.f - z.z) * hv.z + z.z * tanh_fast(v[4 * q + 2]);
            v[4 * q + 3] = (1.f - z.w) * hv.w + z.w * tanh_fast(v[4 * q + 3]);
          }
          store_f32(v, n);
        } else if (EC == EC_PLAIN) {
          if (epi == RNC_EPI_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (epi == RNC_EPI_SIGMOID) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = sigmoid_fast(v[j]);
          }
        } else if (EC != EC_MISC) {
          // (not reached: the z|r class left the chunk above)
        } else if (epi == RNC_EPI_RELU_FLOW) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          if (n <= p.cout && p.cout < n + 32 && valid) {
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
        } else if (epi == RNC_EPI_RELU_ADD_RELU) {
          // residual block tail (extractor.py:55): relu(x + relu(norm(conv(.)))) with the norm folded into the weights
          const float4* rp = reinterpret_cast<const float4*>(p.res + pix * p.ldres + n);
          float4 rreg[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) rreg[q] = valid ? __ldg(rp + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            v[4 * q + 0] = fmaxf(rreg[q].x + fmaxf(v[4 * q + 0], 0.f), 0.f);
            v[4 * q + 1] = fmaxf(rreg[q].y + fmaxf(v[4 * q + 1], 0.f), 0.f);
            v[4 * q + 2] = fmaxf(rreg[q].z + fmaxf(v[4 * q + 2], 0.f), 0.f);
            v[4 * q + 3] = fmaxf(rreg[q].w + fmaxf(v[4 * q + 3], 0.f), 0.f);
          }
        } else if (epi == RNC_EPI_TANH_RELU) {
          // context encoder head (raft_nc_dbl.py:138-140): first half tanh -> net, second half relu -> inp
          if (n < (p.cout >> 1)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = tanh_fast(v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            want_f32 = false;
          }
        }
        if (want_f32) {
          if (p.out_blocked) {
            if (valid) {
              float* dst = p.out_f32 + (static_cast<size_t>(tile) * p.ldo_f32 + n) * kBM + ml;
#pragma unroll
              for (int j = 0; j < 32; ++j) dst[static_cast<size_t>(j) * kBM] = v[j];
            }
          } else {
            store_f32(v, n);
          }
        }
        if (p.out_hi) store_split(v, n);
      }
      // this warp has finished reading the accumulator buffer (tcgen05.wait::ld inside tmem_ld32)
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_cluster(acc_empty_l + buf * 8); else mbar_arrive(&acc_empty[buf]); }
    }
    if (use_stats) flush_stats();
    if (lane == 0) bulk_wait_all0();               // the staging buffer must outlive its TMA stores; their writes complete here
  }

  // ------------------------------------------------------------------ teardown
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();        // the peer may still be multicasting to this CTA's barriers / reading its weight half
  if (warp == 1) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, C::kTmemCols); else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------- host side
// activation plane [B][Hin][Win][ld] halves: 4-D map {C, Win, Hin, B}; the box spans bw x bh input elements and is
// traversed with element strides (sx, sy) -> (bw/sx) x (bh/sy) rows of 64 channels in shared memory
static bool make_in_map(CUtensorMap* m, const void* base, int C, int ld, int B, int Hin, int Win, int bw, int bh, int sx, int sy, bool tf32,
                        long long row_pitch = 0) {
  const cuuint64_t esz = tf32 ? 4 : 2;                        // 128-byte rows: 32 fp32 words or 64 halves
  const cuuint64_t pitch = row_pitch > 0 ? (cuuint64_t)row_pitch : (cuuint64_t)Win * ld;     // elements between image rows
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)B};
  // window view (RNC_CONV_WINDOW): ld < C, i.e. consecutive positions overlap -- the TMA unit walks plain strides
  const cuuint64_t strides[3] = {(cuuint64_t)ld * esz, pitch * esz, (cuuint64_t)Hin * pitch * esz};
  const cuuint32_t box[4] = {tf32 ? 32u : 64u, (cuuint32_t)(bw * sx), (cuuint32_t)(bh * sy), 1};
  const cuuint32_t es[4] = {1, (cuuint32_t)sx, (cuuint32_t)sy, 1};
  return encode_fn()(m, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// weight plane [CoutPad][Ktot] halves: 2-D map {Ktot, CoutPad}, box {64, BN}
static bool make_w_map(CUtensorMap* m, const void* base, int ktot, int coutpad, int bn, bool tf32) {
  const cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)coutpad};
  const cuuint64_t strides[1] = {(cuuint64_t)ktot * (tf32 ? 4 : 2)};
  const cuuint32_t box[2] = {tf32 ? 32u : 64u, (cuuint32_t)bn};
  const cuuint32_t es[2] = {1, 1};
  return encode_fn()(m, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// output plane [B][H][W][ld] (halves or fp32): 4-D map {C, W, H, B}, box = 32 channels x one lane group's bw x bh pixels,
// rows swizzled (64B for halves, 128B for fp32) to match the epilogue's conflict-free staging writes
static bool make_out_map(CUtensorMap* m, const void* base, int C, int ld, int B, int H, int W, int bw, int bh, bool f32) {
  const cuuint64_t esz = f32 ? 4 : 2;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)ld * esz, (cuuint64_t)W * ld * esz, (cuuint64_t)H * W * ld * esz};
  const cuuint32_t box[4] = {32u, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  return encode_fn()(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BN, bool PAIR, int EC>
static int launch(const CUtensorMap* maps, Params& p, cudaStream_t stream, int max_sa, int max_sb) {
  using C = Cfg<BN>;
  const int a_stage = 2 * p.a_plane;
  // ring depths within kRingBudget: both rings hide the same TMA latency, so deepen A (up to 4) while B keeps >= 3 stages
  // a pair's CTA stages half of the weight rows; narrow pair tiles: a whole plane + half of w_hi (see the kernel)
  const int budget = kRingBudget, b_stage = (PAIR && BN <= 64) ? C::kBTile * 3 / 2 : PAIR ? C::kBTile : 2 * C::kBTile;
  const int sa_cap = max_sa > 0 && max_sa < kMaxSA ? max_sa : kMaxSA;      // debug knobs (rnc_conv_umma_desc.flags bits 8-15)
  p.SA = 2;
  while (p.SA < sa_cap && budget - (p.SA + 1) * a_stage >= 3 * b_stage) ++p.SA;
  int sb = (budget - p.SA * a_stage) / b_stage;
  p.SB = sb > 8 ? 8 : sb < 2 ? 2 : sb;
  if (max_sb > 0 && p.SB > max_sb) p.SB = max_sb;
  // Small layers (64 -> 64 3x3: 9 stages of 16 KB): keep the whole weight matrix in the ring for the CTA's lifetime instead
  // of re-streaming it for every pixel tile (the shared-memory fill bandwidth is what bounds narrow tiles).
  p.resident_b = 0;
  {
    const int nb = p.kh * p.kw * p.nblk;
    const int hard = 227 * 1024 - 1024 - 512 - 8192 - 1024 - kStageBytes;
    if (p.ntn == 1 && nb <= kMaxSB && nb > p.SB && max_sb == 0 && 2 * a_stage + nb * b_stage <= hard) {
      p.resident_b = 1; p.SB = nb; p.SA = (hard - nb * b_stage) / a_stage;
      if (p.SA > kMaxSA) p.SA = kMaxSA;
    }
  }
  // + alignment slack, barriers, fp64 statistics slots, 1024-aligned epilogue staging
  const int smem = p.SA * a_stage + p.SB * b_stage + 1024 + 512 + (BN <= 128 ? 8192 : 0) + 1024 + kStageBytes;
  if (smem > 227 * 1024) return RNC_ERR_UNSUPPORTED;
  static unsigned long long done = 0;
  if (int st = ensure_dyn_smem(conv_umma_kernel<BN, PAIR, EC>, 227 * 1024, &done)) return st;
  const int items = (PAIR ? (p.ntiles + 1) / 2 : p.ntiles) * p.ntn;
  const int slots = PAIR ? sm_count() / 2 : sm_count();
  const int grid = (items < slots ? items : slots) * (PAIR ? 2 : 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_umma_kernel<BN, PAIR, EC>, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7],
                                     maps[8], p);
  if (e != cudaSuccess) { g_last_cuda_error = static_cast<int>(e); return RNC_ERR_CUDA; }
  return after_launch();
}

template <bool PAIR, int EC>
static int launch_bn(int bn, const CUtensorMap* maps, Params& p, cudaStream_t s, int msa, int msb) {
  switch (bn) {
    case 32: return launch<32, PAIR, EC>(maps, p, s, msa, msb);
    case 64: return launch<64, PAIR, EC>(maps, p, s, msa, msb);
    case 128: return launch<128, PAIR, EC>(maps, p, s, msa, msb);
    case 192: return launch<192, PAIR, EC>(maps, p, s, msa, msb);
    default: return launch<256, PAIR, EC>(maps, p, s, msa, msb);
  }
}

// CTA-pair (cta_group::2) form of the convolution: the default; RNC_CONV_PAIR=0 selects the single-CTA form
static bool pair_enabled() {
  static const int on = [] { const char* e = getenv("RNC_CONV_PAIR"); return (e && e[0] == '0') ? 0 : 1; }();
  return on != 0;
}

}  // namespace umma
}  // namespace rnc

using namespace rnc;

// Pixel-tile shape of a layer (the A-staging mode decides it): row halo -> 128x1, column halo -> 16x8, per-tap -> TW x 128/TW.
static void tile_shape(int kh, int kw, int stride, int H, int W, int flags, int& TW, int& TH) {
  if (stride == 1 && (flags & RNC_CONV_NO_HALO) == 0 && kw > 1 && W > 64) { TW = 128; TH = 1; }
  else if (stride == 1 && (flags & RNC_CONV_NO_HALO) == 0 && kw == 1 && kh > 1 && W >= 16 && H >= 8) { TW = 16; TH = 8; }
  else { TW = 8; while (TW < W && TW < umma::kBM) TW <<= 1; TH = umma::kBM / TW; }
}

extern "C" long long rnc_conv_umma_tiles(int kh, int kw, int stride, int B, int H, int W, int flags) {
  if (kh <= 0 || kw <= 0 || stride <= 0 || B <= 0 || H <= 0 || W <= 0) return 0;
  int TW, TH;
  tile_shape(kh, kw, stride, H, W, flags, TW, TH);
  return static_cast<long long>(B) * ((W + TW - 1) / TW) * ((H + TH - 1) / TH);
}

// Column-tile width.  The widest tile that divides coutpad is the default; a narrower one (more, smaller work items per
// pixel tile) wins when the grid is under-filled (small batches: 56 pixel tiles on 148 SMs), when the tile count just
// misses a multiple of the SM count (448 tiles = 4 rounds of 256 columns but 7 half-size rounds), or when K is so short
// that the single-buffered 256-column epilogue dominates (1x1 layers).  Cost model in K-step units, fitted to B200
// measurements (tools/bn_probe.py): per K-step 1.0 / 0.77 / 0.53 / 0.36 / 0.30 for 256 / 192 / 128 / 64 / 32 columns, + 3
// K-steps of exposed epilogue for tiles wider than 128 columns (one TMEM accumulator pair), + 0.5 per item, x1.08 when the
// A tile is loaded more than once.
static int choose_bn(const rnc_conv_umma_desc& d, int bn_max, int stride) {
  if (d.flags & RNC_CONV_SPLIT_N) return bn_max == 256 ? 128 : bn_max;
  const int bkc = (d.flags & RNC_CONV_TF32) ? 32 : 64;
  const int taps = d.kh * d.kw, ksteps = taps * ((d.c0 + bkc - 1) / bkc + (d.c1 + bkc - 1) / bkc);
  int TW, TH;
  if (stride == 1 && (d.flags & RNC_CONV_NO_HALO) == 0 && d.kw > 1 && d.W > 64) { TW = 128; TH = 1; }
  else if (stride == 1 && (d.flags & RNC_CONV_NO_HALO) == 0 && d.kw == 1 && d.kh > 1 && d.W >= 16 && d.H >= 8) { TW = 16; TH = 8; }
  else { TW = 8; while (TW < d.W && TW < umma::kBM) TW <<= 1; TH = umma::kBM / TW; }
  const long ntiles = static_cast<long>(d.B) * ((d.W + TW - 1) / TW) * ((d.H + TH - 1) / TH);
  // CTA-pair form (fitted to the same probe): the pair's leader waits for both CTAs' epilogues, so a single-buffered
  // (> 128 columns) tile exposes ~12 K-steps per item, while narrow tiles no longer sit on the shared-memory bandwidth
  const bool pair = umma::pair_enabled() && (d.flags & RNC_CONV_NO_PAIR) == 0;
  static const int cand[5] = {256, 192, 128, 64, 32};
  static const float per_k1[5] = {1.0f, 0.77f, 0.53f, 0.36f, 0.30f}, per_k2[5] = {1.0f, 0.8f, 0.6f, 0.40f, 0.34f};
  const float* per_k = pair ? per_k2 : per_k1;
  const float expose = pair ? 12.0f : 3.0f, per_item = pair ? 1.2f : 0.5f;
  const long slots = pair ? umma::sm_count() / 2 : umma::sm_count();
  const long ptiles = pair ? (ntiles + 1) / 2 : ntiles;
  int best = bn_max;
  float best_cost = 1e30f;
  for (int i = 0; i < 5; ++i) {
    const int bn = cand[i];
    if (bn > bn_max || d.coutpad % bn != 0) continue;
    if ((d.epilogue == RNC_EPI_TANH_RELU || d.epilogue == RNC_EPI_GRU_ZR) && bn < 64) continue;
    const int ntn = d.coutpad / bn;
    const long rounds = (ptiles * ntn + slots - 1) / slots;
    float item = ksteps * per_k[i] + per_item + (bn > 128 ? expose * bn / 256.0f : 0.f);
    if (ntn > 1) item *= 1.08f;
    const float cost = rounds * item;
    if (cost < best_cost * 0.97f) { best_cost = cost; best = bn; }     // wider wins near-ties
  }
  return best;
}

extern "C" int rnc_conv2d_umma_fwd(const rnc_conv_umma_desc* desc, void* stream) {
  using namespace rnc::umma;
  if (!desc) return RNC_ERR_BAD_POINTER;
  const rnc_conv_umma_desc& d = *desc;
  const int stride = d.stride <= 0 ? 1 : d.stride;
  const int Hin = d.hin > 0 ? d.hin : d.H, Win = d.win > 0 ? d.win : d.W;
  if (d.B <= 0 || d.H <= 0 || d.W <= 0 || d.cout <= 0 || d.c0 <= 0 || d.c1 < 0 || stride > 2) return RNC_ERR_BAD_SHAPE;
  if (d.kh < 1 || d.kw < 1 || !(d.kh & 1) || !(d.kw & 1) || d.kh * d.kw > 49) return RNC_ERR_BAD_SHAPE;
  const bool tf32 = (d.flags & RNC_CONV_TF32) != 0;
  const int bk = tf32 ? 32 : kBK, ldmask = tf32 ? 3 : 7;      // operand rows are 16-byte multiples
  if (tf32 && (d.out_hi || d.epilogue == RNC_EPI_RELU_FLOW || d.epilogue == RNC_EPI_GRU_ZR || d.epilogue == RNC_EPI_TANH_RELU))
    return RNC_ERR_UNSUPPORTED;                                // TF32 layers write fp32 outputs only
  const bool window = (d.flags & RNC_CONV_WINDOW) != 0;
  if (window && (tf32 || d.c1 > 0 || d.kw != 1 || stride != 2 || d.win <= 0 || d.hin <= 0 || d.win_pitch < d.win * d.ld0 ||
                 (d.win_pitch & 7)))
    return RNC_ERR_BAD_SHAPE;
  if ((d.ld0 & ldmask) || (d.ld0 < d.c0 && !window) || (d.c1 > 0 && ((d.c0 % bk) != 0 || (d.ld1 & ldmask) || d.ld1 < d.c1))) return RNC_ERR_BAD_SHAPE;
  if (!d.in0_hi || !d.in0_lo || (d.c1 > 0 && (!d.in1_hi || !d.in1_lo)) || !d.w_hi || !d.w_lo || !d.bias) return RNC_ERR_BAD_POINTER;
  if (!aligned16(d.in0_hi) || !aligned16(d.in0_lo) || !aligned16(d.w_hi) || !aligned16(d.w_lo) || !aligned16(d.bias)) return RNC_ERR_BAD_POINTER;
  if (d.c1 > 0 && (!aligned16(d.in1_hi) || !aligned16(d.in1_lo))) return RNC_ERR_BAD_POINTER;
  const int nblk0 = (d.c0 + bk - 1) / bk, nblk1 = (d.c1 + bk - 1) / bk, nblk = nblk0 + nblk1;
  const int ntaps = d.kh * d.kw;
  if (d.ktot != ntaps * nblk * bk) return RNC_ERR_BAD_SHAPE;           // weight planes are [coutpad][taps * blocks * bk]
  int bn;
  if (d.coutpad <= 32) bn = 32; else if (d.coutpad <= 64) bn = 64; else if (d.coutpad <= 128) bn = 128;
  else if (d.coutpad % 192 == 0) bn = 192; else bn = 256;
  // a single CTA stages whole weight tiles: two 64 KB stages of a 256-column tile no longer fit beside the epilogue staging
  if (bn == 256 && !(umma::pair_enabled() && (d.flags & RNC_CONV_NO_PAIR) == 0)) bn = 128;
  if (d.coutpad % bn == 0 && d.epilogue != RNC_EPI_RELU_FLOW && d.epilogue != RNC_EPI_FLOW_DELTA)
    bn = choose_bn(d, bn, stride);
  if (d.coutpad % bn != 0 || d.coutpad < d.cout) return RNC_ERR_BAD_SHAPE;
  if (d.epilogue == RNC_EPI_RELU_FLOW && (d.coutpad < d.cout + 2 || !d.aux0 || !d.out_hi)) return RNC_ERR_BAD_SHAPE;
  if (d.out_hi && (!d.out_lo || (d.ldo_split & 7) || !aligned16(d.out_hi) || !aligned16(d.out_lo))) return RNC_ERR_BAD_POINTER;
  if (d.out_f32 && ((d.ldo_f32 & 3) || !aligned16(d.out_f32))) return RNC_ERR_BAD_POINTER;
  if (d.add && ((d.ldadd & 3) || d.ldadd < d.coutpad || !aligned16(d.add))) return RNC_ERR_BAD_POINTER;
  if (d.stats && (d.epilogue != RNC_EPI_LINEAR || !d.out_f32 || d.out_hi || d.coutpad > 128)) return RNC_ERR_UNSUPPORTED;
  switch (d.epilogue) {
    case RNC_EPI_RELU_ADD_RELU:
      if (!d.res || (d.ldres & 3) || !aligned16(d.res)) return RNC_ERR_BAD_POINTER;
      /* fall through */
    case RNC_EPI_LINEAR: case RNC_EPI_RELU: case RNC_EPI_SIGMOID: case RNC_EPI_RELU_FLOW: case RNC_EPI_TANH_RELU:
      if (!d.out_f32 && !d.out_hi) return RNC_ERR_BAD_POINTER;
      if ((d.cout % 32) != 0 && d.epilogue != RNC_EPI_RELU_FLOW) {
        // the epilogue stores whole 32-channel chunks: the destination must have room for the padded tail
        const int cpad = (d.cout + 31) / 32 * 32;
        if ((d.out_f32 && d.ldo_f32 < cpad) || (d.out_hi && d.ldo_split < cpad)) return RNC_ERR_BAD_SHAPE;
      }
      if (d.epilogue == RNC_EPI_TANH_RELU && (d.cout % 64) != 0) return RNC_ERR_BAD_SHAPE;
      break;
    case RNC_EPI_FLOW_DELTA:
      if (!d.aux0 || d.cout != 2 || d.out_hi) return RNC_ERR_BAD_SHAPE;
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

  // ---- tile shape and A-staging mode
  Params p;
  p.B = d.B; p.H = d.H; p.W = d.W;
  p.kh = d.kh; p.kw = d.kw; p.ph = d.kh / 2; p.pw = d.kw / 2; p.stride = stride; p.sx = window ? 1 : stride;
  int mode = MODE_TAP;
  if (stride == 1 && (d.flags & RNC_CONV_NO_HALO) == 0) {
    if (d.kw > 1 && d.W > 64) mode = MODE_ROWHALO;                 // one image row x 128 px per tile
    else if (d.kw == 1 && d.kh > 1 && d.W >= 16 && d.H >= 8) mode = MODE_COLHALO;
  }
  int TW, TH, box_w, box_h;
  if (mode == MODE_ROWHALO) { TW = 128; TH = 1; box_w = 136; box_h = 1; }
  else if (mode == MODE_COLHALO) { TW = 16; TH = 8; box_w = 16; box_h = TH + 2 * p.ph; }
  else {
    TW = 8;
    while (TW < d.W && TW < kBM) TW <<= 1;
    TH = kBM / TW; box_w = TW; box_h = TH;
  }
  if (box_w * box_h > 256 - 8 && mode == MODE_COLHALO) return RNC_ERR_UNSUPPORTED;   // kh <= 9
  p.mode = mode; p.TW = TW; p.TH = TH;
  {
    static const char* env = getenv("RNC_CONV_PROBE_NOB");
    p.probe_nob = env != nullptr && env[0] == '1';
  }
  p.tiles_x = (d.W + TW - 1) / TW; p.tiles_y = (d.H + TH - 1) / TH;
  p.ntiles = d.B * p.tiles_x * p.tiles_y; p.ntn = d.coutpad / bn;
  p.a_plane = box_w * box_h * 128;
  p.nblk0 = nblk0; p.nblk = nblk; p.bk = bk; p.tf32 = tf32 ? 1 : 0;
  p.cout = d.cout; p.epilogue = d.epilogue; p.unscale = d.unscale; p.bias = d.bias;
  p.out_f32 = d.out_f32; p.ldo_f32 = d.ldo_f32;
  p.out_hi = static_cast<__half*>(d.out_hi); p.out_lo = static_cast<__half*>(d.out_lo); p.ldo_split = d.ldo_split;
  p.h = d.h; p.ldh = d.ldh; p.aux0 = d.aux0; p.ldaux = d.ldaux; p.res = d.res; p.ldres = d.ldres;
  p.stats = d.stats;
  p.add = d.add; p.ldadd = d.ldadd;
  p.aux_blocked = (d.flags & RNC_CONV_AUX_BLOCKED) ? 1 : 0;
  p.out_blocked = (d.flags & RNC_CONV_OUT_BLOCKED) ? 1 : 0;
  if (p.out_blocked && (d.epilogue != RNC_EPI_LINEAR || d.out_hi || d.stats)) return RNC_ERR_UNSUPPORTED;
  if (p.aux_blocked && d.epilogue != RNC_EPI_GRU_ZR && d.epilogue != RNC_EPI_GRU_Q) return RNC_ERR_UNSUPPORTED;

  CUtensorMap maps[9];
  const long long pitch0 = window ? d.win_pitch : 0;
  bool ok = make_in_map(&maps[0], d.in0_hi, d.c0, d.ld0, d.B, Hin, Win, box_w, box_h, p.sx, stride, tf32, pitch0) &&
            make_in_map(&maps[1], d.in0_lo, d.c0, d.ld0, d.B, Hin, Win, box_w, box_h, p.sx, stride, tf32, pitch0);
  if (d.c1 > 0) {
    ok = ok && make_in_map(&maps[2], d.in1_hi, d.c1, d.ld1, d.B, Hin, Win, box_w, box_h, stride, stride, tf32) &&
         make_in_map(&maps[3], d.in1_lo, d.c1, d.ld1, d.B, Hin, Win, box_w, box_h, stride, stride, tf32);
  } else {
    maps[2] = maps[0]; maps[3] = maps[1];
  }
  const bool pair = umma::pair_enabled() && (d.flags & RNC_CONV_NO_PAIR) == 0;
  ok = ok && make_w_map(&maps[4], d.w_hi, d.ktot, d.coutpad, pair ? bn / 2 : bn, tf32) &&
       make_w_map(&maps[5], d.w_lo, d.ktot, d.coutpad, pair ? bn / 2 : bn, tf32);
  // epilogue stores: out_hi / out_lo (and h for the q gate) as boxes of 32 channels x one TMEM lane group's pixels
  {
    const int obw = TW < 32 ? TW : 32, obh = 32 / obw;
    const int flow2 = d.epilogue == RNC_EPI_RELU_FLOW ? 2 : 0;
    const int cw = d.epilogue == RNC_EPI_GRU_ZR ? d.cout / 2 : (d.cout + flow2 + 31) / 32 * 32;
    if (d.out_hi) {
      if ((d.ldo_split & 7) || !aligned16(d.out_hi) || !aligned16(d.out_lo) || d.ldo_split < cw) return RNC_ERR_BAD_POINTER;
      ok = ok && make_out_map(&maps[6], d.out_hi, cw, d.ldo_split, d.B, d.H, d.W, obw, obh, false) &&
           make_out_map(&maps[7], d.out_lo, cw, d.ldo_split, d.B, d.H, d.W, obw, obh, false);
    } else {
      maps[6] = maps[0]; maps[7] = maps[0];
    }
    if (d.epilogue == RNC_EPI_GRU_Q) {
      if (!aligned16(d.h) || d.ldh < d.cout || d.out_f32) return RNC_ERR_BAD_POINTER;
      ok = ok && make_out_map(&maps[8], d.h, d.cout, d.ldh, d.B, d.H, d.W, obw, obh, true);
    } else if (d.out_f32 && !p.out_blocked && d.epilogue != RNC_EPI_FLOW_DELTA) {
      const int cf = d.epilogue == RNC_EPI_TANH_RELU ? d.cout / 2 : (d.cout + 31) / 32 * 32;   // TANH_RELU: only the tanh half
      if (d.ldo_f32 < cf) return RNC_ERR_BAD_SHAPE;
      ok = ok && make_out_map(&maps[8], d.out_f32, cf, d.ldo_f32, d.B, d.H, d.W, obw, obh, true);
    } else {
      maps[8] = maps[0];
    }
  }
  if (!ok) return RNC_ERR_BAD_SHAPE;

  cudaStream_t s = as_stream(stream);
  const int msa = (d.flags >> 8) & 15, msb = (d.flags >> 12) & 15;
  const bool plain = (d.epilogue == RNC_EPI_LINEAR || d.epilogue == RNC_EPI_RELU || d.epilogue == RNC_EPI_SIGMOID) && !d.stats;
  const int ec = d.epilogue == RNC_EPI_GRU_ZR ? EC_GRU_ZR : d.epilogue == RNC_EPI_GRU_Q ? EC_GRU_Q : plain ? EC_PLAIN : EC_MISC;
  if (pair) {
    switch (ec) {
      case EC_GRU_ZR: return launch_bn<true, EC_GRU_ZR>(bn, maps, p, s, msa, msb);
      case EC_GRU_Q: return launch_bn<true, EC_GRU_Q>(bn, maps, p, s, msa, msb);
      case EC_MISC: return launch_bn<true, EC_MISC>(bn, maps, p, s, msa, msb);
      default: return launch_bn<true, EC_PLAIN>(bn, maps, p, s, msa, msb);
    }
  }
  switch (ec) {
    case EC_GRU_ZR: return launch_bn<false, EC_GRU_ZR>(bn, maps, p, s, msa, msb);
    case EC_GRU_Q: return launch_bn<false, EC_GRU_Q>(bn, maps, p, s, msa, msb);
    case EC_MISC: return launch_bn<false, EC_MISC>(bn, maps, p, s, msa, msb);
    default: return launch_bn<false, EC_PLAIN>(bn, maps, p, s, msa, msb);
  }
}
