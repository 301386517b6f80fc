// Implicit-GEMM convolution / GEMM for sm_100a: TMA-fed tcgen05.mma with TMEM accumulators.
//
// One persistent, warp-specialised kernel serves every dense contraction on the DiffusionVID hot path:
//   * backbone convolutions (detectron2 ResNet-101 + FPN restated in SURVEY.md A1; FrozenBN folded into weight/bias),
//   * every nn.Linear of the decoder (reference mega_core/modeling/roi_heads/box_head/box_head.py:447-491,675-684).
// Activations are NHWC fp16, weights [Cout][R*S*Cin] fp16 (K-major), accumulation fp32 in TMEM.
// A GEMM [M,K]x[N,K]^T is the 1x1 "convolution" over an image of height 1 and width M.
//
// Tile: 128 output pixels (th x tw patch of one image) x BN output channels, K step 64 channels of one filter tap.
// The A tile of tap (r,s) is the same 4-D TMA box shifted by (r-pad, s-pad); out-of-bounds rows are zero-filled by
// TMA, which implements the convolution's zero padding and all tile tails without any predicate in the kernel.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = TMA store issuer, 4..11 = epilogue (TMEM -> regs
// -> swizzled smem, two groups alternating 64-column chunks). Two TMEM accumulator stages let the epilogue of tile i
// overlap the main loop of tile i+1.
#include <stdlib.h>
#include "ptx_sm100.cuh"
#include "dvid_internal.h"

namespace dvid {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int OUT_STAGE_BYTES = BLOCK_M * 64 * 2;     // 16 KB, two of them (ping-pong)

// n / d for 0 <= n < 2^31 and a divisor fixed on the host: q = umulhi(n, mul) >> shr  (d == 1: q = n)
struct FastDiv {
  unsigned d, mul, shr;
  __host__ void init(int denom) {
    d = static_cast<unsigned>(denom < 1 ? 1 : denom);
    if (d == 1) { mul = 0; shr = 0; return; }
    unsigned lg = 0;
    while ((1ull << lg) < d) ++lg;                 // ceil(log2(d))
    const unsigned pw = 31 + lg;
    mul = static_cast<unsigned>(((1ull << pw) + d - 1) / d);
    shr = pw - 32;
  }
  __device__ __forceinline__ int div(int n) const {
    return d == 1 ? n : static_cast<int>(__umulhi(static_cast<unsigned>(n), mul) >> shr);
  }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const {
    q = div(n);
    r = n - q * static_cast<int>(d);
  }
};

struct ConvGemmParams {
  int n_img, h_out, w_out;
  int cin, cout;
  int R, S, pad, stride;
  int tw_log2, th;  // tile = th x (1 << tw_log2) pixels, th * tw == 128
  int tiles_x, tiles_y;
  int m_tiles, n_tiles;
  int kb_per_tap;    // ceil(cin / 64)
  int total_kb;      // R * S * kb_per_tap
  int splits;        // split-K factor (fp32 partial output only)
  int kb_per_split;  // ceil(total_kb / splits)
  const float* bias;     // [cout] or nullptr
  const __half* resid;   // NHWC [n_img, resid_h, resid_w, cout] or nullptr; read at (y >> resid_shift, x >> resid_shift)
  int resid_shift, resid_h, resid_w;
  int relu;              // activation: 0 none, 1 ReLU, 2 GELU(erf)
  float* out_f32;        // if set: fp32 output [splits][M][cout] by direct stores (GEMM mode), no bias/resid/relu
  long long m_total;     // rows of the GEMM view (n_img * h_out * w_out)
  int dbg;               // DVID_DBG experiment bits (timing experiments only; results are wrong when set)
  // division by the (launch-invariant) tile counts as multiply-high + shift: every role decodes its tile index once per
  // tile and a 32-bit integer division is a ~35-instruction dependent chain - five of them cost the epilogue ~0.4 us
  // per tile (device trace), as much as converting a 64-column chunk
  FastDiv div_m_tiles, div_n_tiles, div_tiles_x, div_tiles_y;
  unsigned long long* trace;   // DVID_TRACE: per-role event log of CTA `trace_cta` (debug only), else nullptr
  int trace_cta;               // DVID_TRACE_CTA (default 0)
  // stream-K (SK variants only): fp32 partial accumulator tiles [grid][BN/4][128][4] and one flag per CTA
  // weights are independent of the previous kernel: every CTA asks L2 for its slice of them BEFORE griddepcontrol.wait,
  // while the previous kernel is still draining (its output, our A operand, is the only thing we have to wait for)
  const uint8_t* w_base;
  unsigned w_slice_bytes;      // 0: no prefetch; else multiple of 128
  unsigned long long w_bytes;
  int b_pre;                   // BSTAT: first resident weight tile requested before the dependency wait (DVID_BPRE)
  float* sk_ws;
  unsigned* sk_flags;
  FastDiv div_total_kb;
  // PATCH variants (3x3 / stride 1): bytes of one A patch ((th + 2) rows x tw pixels x 128 B) and of one patch row
  int patch_bytes, patch_row_bytes;
};

// BSTAT ("B-stationary", K <= 256): the whole [BN x K] weight tile stays resident in smem while the CTA walks a
// contiguous range of M tiles (m fastest), so the weights are read from L2 once per CTA instead of once per tile and
// only the A tiles stream through the ring.  L2->SM bandwidth (~6.5 TB/s for unique lines) is what bounds the short-K
// GEMMs (dynamic_layer 2400x256->32768: 467 MB of operand reads at 128x256 tiles -> ~190 MB).
constexpr int BSTAT_MAX_KB = 4;   // K <= 256
// named barriers: 1 = the four epilogue warps; FULL/FREE = epilogue warps <-> TMA store warp, per staging buffer
constexpr int BAR_FULL0 = 2, BAR_FREE0 = 4, BAR_BIAS0 = 6;

// RES (residual with resid_shift == 0): the residual tile is ADDED BY THE TENSOR CORE.  Each 64-channel chunk of the
// residual (same 128-pixel patch as the output tile) is TMA-loaded into an A-ring stage like an extra k-block and
// multiplied by a 16x16 identity held in smem: D[:, 64j+16k .. +16] += R_j[:, 16k .. +16] * I (four N=16 MMAs per
// chunk, fp16 x 1.0 accumulated in fp32 = exact).  The first version added it in the epilogue from per-thread strided
// 16-byte global loads, which cost 3-4 us per 64-column chunk (device trace: profiles/r01_trace_res4_conv3.txt).
constexpr int IDENT_BYTES = 16 * 128;   // 16 rows (n) x 64 k fp16, 128-byte swizzle; I[n][k] = (n == k), k < 16

// SK ("stream-K"): the (tile, k-block) space is cut into gridDim.x equal contiguous ranges instead of whole tiles, so a
// layer of 152 tiles costs 152/148 of a wave instead of two (res4: 8 frames x 38x64 pixels = 152 M tiles, 46 of the
// 104 backbone convs).  With tiles >= CTAs every range is at least one tile long, so a tile is shared by at most two
// CTAs: the range of CTA b ends with the HEAD k-blocks of a tile whose TAIL k-blocks open the range of CTA b+1.  A CTA
// runs its head segment FIRST and parks the fp32 accumulator tile in its workspace slot (coalesced: 4 columns x 128
// rows per 2 KB) and raises a flag; the tail owner adds the parked tile in its epilogue and finishes the tile (bias /
// residual / activation / TMA store) as usual.  The parked tile is the producer's first item and depends on nothing,
// CTAs are dispatched in index order and are co-resident (1 per SM, grid <= SMs): no deadlock, and both sides of the
// hand-off overlap a main loop (the first version parked tails and finished heads as each CTA's LAST item: the
// exposed wait + 128 KB read cost as much as the saved second wave).
// PATCH ("column patch", 3x3 stride-1 convolutions): the three taps of one filter COLUMN read the same pixels shifted by
// whole image rows, so one TMA box of (th + 2) rows x tw pixels serves three k-blocks - tap (dy, dx) is the view that
// starts dy rows (dy * tw * 128 B, a multiple of the 1024-byte swizzle atom for tw >= 8) into the patch of column dx.
// The A operand then crosses L2->SM 3 * (th + 2) / th times per channel block instead of 9 times (th = 8: 3.75), which
// is what bounds the main loop of these layers (a 64-wide tile costs 0.55 of a 256-wide one: bytes per k-block, not
// MMA time).  A patches and B tiles ride separate rings (PA x 24 KB, PB x B_STAGE_BYTES); the k-blocks of a tile are
// visited column by column: (dx, channel block, dy).
constexpr int A_PATCH_BYTES = 24 * 1024;   // (th + 2) * tw * 128 <= 24 KB: (th, tw) in {(4,32), (8,16), (16,8)}

// CTA2 ("CTA pair", BN = 256 plain convolutions): the two CTAs of a {2,1,1} cluster compute two adjacent M tiles with ONE
// tcgen05.mma.cta_group::2 of M = 256 per k-step.  Each CTA loads its own A tile (16 KB per k-block) and only HALF of the
// B tile (16 KB instead of 32 KB): what bounds these layers is the bytes an SM can take in per k-block (measured: every
// tile width runs at ~90 GB/s per SM), so 32 KB instead of 48 KB per k-block and a six-deep ring instead of four.
template <int BN, bool BSTAT = false, bool RES = false, bool PATCH = false, bool CTA2 = false>
struct ConvGemmCfg {
  static constexpr int B_STAGE_BYTES = (CTA2 ? BN / 2 : BN) * BLOCK_K * 2;
  static constexpr int STAGES0 = CTA2 ? 6 : (BSTAT ? ((BN == 256) ? 4 : 8) : ((BN == 256) ? 4 : ((BN == 128) ? 6 : 8)));
  static constexpr int STAGES = STAGES0 - (RES ? 1 : 0);   // room for the identity tile (and keeps BN=256 under 227 KB)
  // ring depth in k-blocks is what the saved A bytes buy: 3 * PA and PB k-blocks in flight (plain walk: 4 / 6 / 8)
  static constexpr int PA = (BN == 256) ? 2 : ((BN == 128) ? 3 : 4);   // A patch stages (24 KB each)
  static constexpr int PB = (BN == 256) ? 4 : ((BN == 128) ? 7 : 12);  // B tile stages
  static constexpr int NBAR = PATCH ? (PA + PB) : STAGES;              // full / empty barrier pairs of the rings
  static constexpr int CTL_BYTES = PATCH ? 384 : 256;                  // mbarriers + TMEM slot
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages
  static constexpr int RING_BYTES =
      PATCH ? (PA * A_PATCH_BYTES + PB * B_STAGE_BYTES)
            : (BSTAT ? (STAGES * A_STAGE_BYTES + BSTAT_MAX_KB * B_STAGE_BYTES) : (STAGES * (A_STAGE_BYTES + B_STAGE_BYTES)));
  static constexpr int SMEM_BYTES = RING_BYTES + 2 * OUT_STAGE_BYTES + 1024 /*align*/ + CTL_BYTES /*barriers*/ +
                                    (BN * 4 > 512 ? BN * 4 : 512) /*bias staging: each epilogue group its own half*/ +
                                    (RES ? IDENT_BYTES : 0);
  static_assert(!PATCH || (!BSTAT && !RES), "the column-patch walk serves the plain variants only");
  static_assert(!CTA2 || (!BSTAT && !RES && !PATCH && BN == 256), "CTA pairs serve the plain 256-wide variant only");
  static_assert(2 * NBAR * 8 + 6 * 8 + 4 <= CTL_BYTES, "barrier block");
};

// GELU, erf form (torch.nn.GELU default; Swin MLP, swintransformer.py:47-66): x * Phi(x).  erff() costs ~30
// instructions per element and made the fc1 epilogue 2x longer than its K=512 mainloop, so Phi is evaluated as
// 2^-g(|x|) with g = -log2(Phi(-u)) a degree-6 minimax polynomial on u in [0,5] (|x| clamped to 5, where Phi(-5) is
// 2.9e-7): 6 FMA + one ex2.  Max abs error 2.9e-6, max relative error 1.5e-5 - far inside the fp16 output rounding.
__device__ __forceinline__ float gelu_erf(float x) {
  const float u = fminf(fabsf(x), 5.f);
  float g = fmaf(-2.975744064e-05f, u, 7.135751075e-04f);
  g = fmaf(g, u, -7.775241509e-03f);
  g = fmaf(g, u, 5.269032717e-02f);
  g = fmaf(g, u, 4.595123231e-01f);
  g = fmaf(g, u, 1.150926590e+00f);
  g = fmaf(g, u, 1.000011683e+00f);
  float h;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(-g));
  return x * (x > 0.f ? 1.f - h : h);
}

// debug event log (DVID_TRACE=1): role r appends (code, clock) pairs to its own 1024-entry lane of p.trace, CTA 0 only
__device__ __forceinline__ void trace_ev(const ConvGemmParams& p, int role, int& n, unsigned code) {
  if (p.trace != nullptr && blockIdx.x == p.trace_cta && n < 1023) {
    const unsigned long long t = static_cast<unsigned long long>(clock64());   // SM-local cycles (cheap)
    p.trace[role * 2048 + 2 * n] = code;
    p.trace[role * 2048 + 2 * n + 1] = t;
    ++n;
  }
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
constexpr int BAR_SK = 8;   // named barrier of the eight epilogue warps (stream-K hand-off)

template <int BN, bool BSTAT, bool RES, bool SK, bool PATCH = false, bool CTA2 = false>
__global__ void __launch_bounds__(384, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                 const ConvGemmParams p) {
  using Cfg = ConvGemmCfg<BN, BSTAT, RES, PATCH, CTA2>;
  static_assert(!CTA2 || !SK, "no stream-K for CTA pairs");
  // residual added in the epilogue registers (FPN top-down adds, 64-wide tiles): not compiled into the 256-wide
  // variants, whose epilogue keeps two chunks of accumulator columns in flight and has no registers to spare for it -
  // the host sends such layers to the 128-wide tile
  constexpr bool ERES = !RES && BN < 256;
  const int rank = CTA2 ? static_cast<int>(cluster_ctarank()) : 0;   // 0 = leader (issues the MMAs of the pair)
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NBAR = Cfg::NBAR;
  constexpr int PA = Cfg::PA, PB = Cfg::PB;
  static_assert(!PATCH || !SK, "no stream-K in the column-patch walk");
  extern __shared__ uint8_t smem_raw[];
  // align to 1024 B (128-byte swizzle atoms) with pointer arithmetic on the __shared__ symbol, so the compiler keeps
  // the shared state space (LDS/STS instead of generic LD/ST in the epilogue)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  // B: ring (normal / PATCH) or the resident [kb][BN x 64] weight tile (BSTAT)
  uint8_t* sB = sA + (PATCH ? PA * A_PATCH_BYTES : STAGES * A_STAGE_BYTES);
  uint8_t* sOut = smem + Cfg::RING_BYTES;
  uint8_t* sIdent = sOut + 2 * OUT_STAGE_BYTES;   // RES: 16x16 identity B tile (1024-byte aligned for the 128B swizzle)
  uint8_t* sCtl = sIdent + (RES ? IDENT_BYTES : 0);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sCtl);
  uint64_t* empty_bar = full_bar + NBAR;          // PATCH: [0, PA) = A patch ring, [PA, PA + PB) = B ring
  uint64_t* tmem_full = empty_bar + NBAR;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* bres_full = tmem_empty + 2;           // BSTAT: resident weights landed / may be overwritten
  uint64_t* bres_empty = bres_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_empty + 1);
  float* sBias = reinterpret_cast<float*>(sCtl + Cfg::CTL_BYTES);   // [BN] bias of the current tile

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_trigger();   // the next kernel of the stream may start its prologue while this one runs (see dvid_internal.h)
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < NBAR; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], CTA2 ? 16 : 8);  // one arrive per epilogue warp (CTA2: of both CTAs, on the leader's)
    }
    mbar_init(bres_full, 1);
    mbar_init(bres_empty, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if (CTA2) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  if (warp == 3 && p.w_slice_bytes != 0 && elect_one()) {
    const unsigned long long off = static_cast<unsigned long long>(blockIdx.x) * p.w_slice_bytes;
    if (off < p.w_bytes) {
      const unsigned long long left = p.w_bytes - off;
      const unsigned n = left < p.w_slice_bytes ? static_cast<unsigned>(left) & ~15u : p.w_slice_bytes;
      if (n != 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.w_base + off), "r"(n) : "memory");
    }
  }
  if (RES && warp >= 4) {
    // identity B tile: row n (16 rows of 128 B), element k at 16-byte chunk (k / 8) ^ (n & 7)
    const int t = threadIdx.x - 128;
    if (t < IDENT_BYTES / 16) {
      const int n = t >> 3, chunk = t & 7;        // logical chunk = 8 k-elements
      uint4 v = make_uint4(0, 0, 0, 0);
      if (chunk == (n >> 3)) {                     // k == n lies in logical chunk n / 8, element n % 8
        uint32_t w[4] = {0, 0, 0, 0};
        w[(n & 7) >> 1] = (n & 1) ? 0x3C000000u : 0x00003C00u;   // fp16 1.0 in the low / high half
        v = make_uint4(w[0], w[1], w[2], w[3]);
      }
      *reinterpret_cast<uint4*>(sIdent + n * 128 + ((chunk ^ (n & 7)) << 4)) = v;
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all();      // the peer's barriers are initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // CTA2: the pair walks PAIRS of adjacent M tiles (an odd tail tile is paired with an out-of-range one: its loads are
  // zero-filled and its stores clipped by TMA)
  const int total_tiles = CTA2 ? ((p.m_tiles + 1) >> 1) * p.n_tiles : p.m_tiles * p.n_tiles * p.splits;
  const int tw = 1 << p.tw_log2;
  // tile walk: normal = strided over the grid, n fastest; BSTAT = one contiguous range per CTA, m fastest (splits == 1)
  const int tile_begin = BSTAT ? static_cast<int>(static_cast<long long>(total_tiles) * blockIdx.x / gridDim.x)
                               : static_cast<int>(CTA2 ? (blockIdx.x >> 1) : blockIdx.x);
  const int tile_end = BSTAT ? static_cast<int>(static_cast<long long>(total_tiles) * (blockIdx.x + 1) / gridDim.x)
                             : total_tiles;
  const int tile_step = BSTAT ? 1 : static_cast<int>(CTA2 ? (gridDim.x >> 1) : gridDim.x);
  auto decode = [&](int tile, int& n_idx, int& m_idx, int& split) {
    if (BSTAT) {
      p.div_m_tiles.divmod(tile, n_idx, m_idx);
      split = 0;
    } else if (CTA2) {
      int pair;
      p.div_n_tiles.divmod(tile, pair, n_idx);
      m_idx = 2 * pair + rank;
      split = 0;
    } else {
      int rest;
      p.div_n_tiles.divmod(tile, rest, n_idx);
      p.div_m_tiles.divmod(rest, split, m_idx);
    }
  };
  // m tile -> (x tile, y tile, image)
  auto locate = [&](int m_idx, int& tx, int& ty, int& img) {
    int rest;
    p.div_tiles_x.divmod(m_idx, rest, tx);
    p.div_tiles_y.divmod(rest, img, ty);
  };
  auto k_range = [&](int split, int& kb_begin, int& kb_end) {
    kb_begin = split * p.kb_per_split;
    kb_end = min(p.total_kb, kb_begin + p.kb_per_split);
  };
  // work walk: every role iterates the same (tile, k-block range) items.  Normal: whole tiles (or split-K slices).
  // SK: this CTA's contiguous range [sk_u0, sk_u1) of the flattened (tile, k-block) space, cut at tile boundaries.
  static_assert(!SK || (!BSTAT && !RES), "stream-K serves the plain variants only");
  int sk_u0 = 0, sk_u1 = 0, sk_uh = 0;
  if (SK) {
    const long long units = static_cast<long long>(total_tiles) * p.total_kb;
    sk_u0 = static_cast<int>(units * blockIdx.x / gridDim.x);
    sk_u1 = static_cast<int>(units * (blockIdx.x + 1) / gridDim.x);
    sk_uh = p.div_total_kb.div(sk_u1) * p.total_kb;     // start of the trailing HEAD segment (== sk_u1: none)
  }
  // SK item order: the trailing head segment FIRST (its accumulator is parked for the next CTA while this CTA's other
  // items run), then the range in order - a leading tail segment (finished with the tile parked by CTA b-1, whose
  // first item that was) and whole tiles.  Every hand-off therefore overlaps a main loop; nothing waits at the end.
  struct Walk { int tile, kb0, kb1, u; bool head; };
  auto walk_set = [&](Walk& w) {
    if (SK) {
      const int limit = w.head ? sk_u1 : sk_uh;
      if (w.u < limit) {
        w.tile = p.div_total_kb.div(w.u);
        w.kb0 = w.u - w.tile * p.total_kb;
        w.kb1 = min(p.total_kb, w.kb0 + (limit - w.u));
      }
    }
  };
  auto walk_begin = [&]() {
    Walk w;
    w.tile = tile_begin; w.kb0 = 0; w.kb1 = 0;
    w.head = SK && sk_uh < sk_u1;
    w.u = w.head ? sk_uh : sk_u0;
    walk_set(w);
    return w;
  };
  auto walk_valid = [&](const Walk& w) { return SK ? (w.head || w.u < sk_uh) : (w.tile < tile_end); };
  auto walk_next = [&](Walk& w) {
    if (SK) {
      if (w.head) { w.head = false; w.u = sk_u0; }
      else w.u += w.kb1 - w.kb0;
      walk_set(w);
    } else {
      w.tile += tile_step;
    }
  };
  auto walk_k = [&](const Walk& w, int split, int& kb_begin, int& kb_end) {
    if (SK) { kb_begin = w.kb0; kb_end = w.kb1; }
    else k_range(split, kb_begin, kb_end);
  };

  // BSTAT: the first resident weight tile (up to 128 KB per CTA, 19 MB over the grid) does not depend on the previous
  // kernel - it is requested before the dependency wait, so it lands while that kernel drains
  int pre_n = -1;
  if (BSTAT && p.b_pre && warp == 0 && tile_begin < tile_end) {
    int n0, m0, s0;
    decode(tile_begin, n0, m0, s0);
    pre_n = n0;
    if (elect_one()) {
      mbar_expect_tx(bres_full, p.total_kb * Cfg::B_STAGE_BYTES);
      for (int kb = 0, tap = 0, cblk = 0; kb < p.total_kb; ++kb) {
        tma_load_2d(sB + kb * Cfg::B_STAGE_BYTES, &tmB, bres_full, tap * p.cin + cblk * BLOCK_K, n0 * BN);
        if (++cblk == p.kb_per_tap) { cblk = 0; ++tap; }
      }
    }
    __syncwarp();
  }
  pdl_wait();      // everything above overlapped the previous kernel's tail; its outputs are visible from here on

  if (warp == 0) {
    if (elect_one()) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      int pstage = 0;            // PATCH: A patch ring position (stage / phase walk the B ring)
      uint32_t pphase = 0;
      int cur_n = pre_n;
      uint32_t bphase = 0;
      int tn = 0;
      for (Walk wk = walk_begin(); walk_valid(wk); walk_next(wk)) {
        const int tile = wk.tile;
        int n_idx, m_idx, split;
        decode(tile, n_idx, m_idx, split);
        int tx, ty, img;
        locate(m_idx, tx, ty, img);
        const int x_in0 = tx * tw * p.stride - p.pad;
        const int y_in0 = ty * p.th * p.stride - p.pad;
        int kb_begin, kb_end;
        walk_k(wk, split, kb_begin, kb_end);
        if (BSTAT && n_idx != cur_n) {
          if (cur_n >= 0) {                       // every MMA that reads the old weights has completed
            mbar_wait(bres_empty, bphase);
            bphase ^= 1;
          }
          cur_n = n_idx;
          mbar_expect_tx(bres_full, p.total_kb * Cfg::B_STAGE_BYTES);
          for (int kb = 0, tap = 0, cblk = 0; kb < p.total_kb; ++kb) {
            tma_load_2d(sB + kb * Cfg::B_STAGE_BYTES, &tmB, bres_full, tap * p.cin + cblk * BLOCK_K, n_idx * BN);
            if (++cblk == p.kb_per_tap) { cblk = 0; ++tap; }
          }
        }
        if (PATCH) {
          // column by column: one patch of (th + 2) rows per (dx, channel block), then the three weight tiles of its taps
          for (int dx = 0; dx < 3; ++dx) {
            for (int cb = 0; cb < p.kb_per_tap; ++cb) {
              mbar_wait(&empty_bar[pstage], pphase ^ 1);
              mbar_expect_tx(&full_bar[pstage], p.patch_bytes);
              tma_load_4d(sA + pstage * A_PATCH_BYTES, &tmA, &full_bar[pstage], cb * BLOCK_K, x_in0 + dx, y_in0, img);
              if (++pstage == PA) { pstage = 0; pphase ^= 1; }
              for (int dy = 0; dy < 3; ++dy) {
                mbar_wait(&empty_bar[PA + stage], phase ^ 1);
                mbar_expect_tx(&full_bar[PA + stage], Cfg::B_STAGE_BYTES);
                tma_load_2d(sB + stage * Cfg::B_STAGE_BYTES, &tmB, &full_bar[PA + stage],
                            (dy * 3 + dx) * p.cin + cb * BLOCK_K, n_idx * BN);
                if (++stage == PB) { stage = 0; phase ^= 1; }
              }
            }
          }
          continue;
        }
        // (tap, channel block, filter row, filter column) of the k-block, advanced incrementally: the divisions
        // they replace were a dependent ~70-instruction chain per k-block on the single producer thread
        int tap = 0, cblk = 0, r = 0, s = 0;
        if (kb_begin != 0) {                       // split-K slices only
          tap = kb_begin / p.kb_per_tap;
          cblk = kb_begin - tap * p.kb_per_tap;
          r = tap / p.S;
          s = tap - r * p.S;
        }
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (!(p.dbg & 256)) trace_ev(p, 0, tn, (tile << 8) | kb);     // slot free, load issued
          if (CTA2) {
            // the leader's barrier counts the bytes of both CTAs; each CTA brings its own A rows and half of the B rows
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + Cfg::B_STAGE_BYTES));
            tma_load_4d_pair(sA + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], cblk * BLOCK_K, x_in0 + s, y_in0 + r,
                             img);
            tma_load_2d_pair(sB + stage * Cfg::B_STAGE_BYTES, &tmB, &full_bar[stage], tap * p.cin + cblk * BLOCK_K,
                             n_idx * BN + rank * (BN / 2));
          } else {
            mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + (BSTAT ? 0 : Cfg::B_STAGE_BYTES));
            tma_load_4d(sA + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], cblk * BLOCK_K, x_in0 + s, y_in0 + r, img);
            if (!BSTAT)
              tma_load_2d(sB + stage * Cfg::B_STAGE_BYTES, &tmB, &full_bar[stage], tap * p.cin + cblk * BLOCK_K,
                          n_idx * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++cblk == p.kb_per_tap) {
            cblk = 0;
            ++tap;
            if (++s == p.S) { s = 0; ++r; }
          }
        }
        if (RES) {
          const int nchunks = min(BN / 64, (p.cout - n_idx * BN + 63) / 64);
          for (int c = 0; c < nchunks; ++c) {      // residual chunks ride the A ring like extra k-blocks
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES);
            tma_load_4d(sA + stage * A_STAGE_BYTES, &tmR, &full_bar[stage], n_idx * BN + c * 64, tx * tw, ty * p.th,
                        img);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (CTA2: of the leader CTA, for the pair) =====================
    constexpr uint32_t idesc = umma_idesc_f16(CTA2 ? 2 * BLOCK_M : BLOCK_M, BN);
    int stage = 0;
    uint32_t phase = 0;
    int pstage = 0;              // PATCH: A patch ring position (stage / phase walk the B ring)
    uint32_t pphase = 0;
    int as = 0;
    uint32_t aphase = 0;
    int cur_n = -1;
    uint32_t bphase = 0;
    int tn = 0;
    for (Walk wk = walk_begin(); walk_valid(wk); walk_next(wk)) {
      const int tile = wk.tile;
      int n_idx, m_idx, split;
      decode(tile, n_idx, m_idx, split);
      int kb_begin, kb_end;
      walk_k(wk, split, kb_begin, kb_end);
      if (BSTAT && n_idx != cur_n) {
        cur_n = n_idx;
        mbar_wait(bres_full, bphase);
        bphase ^= 1;
      }
      bool last_of_n = false;
      if (BSTAT) {
        const int nt = tile + tile_step;
        last_of_n = (nt < tile_end) && (p.div_m_tiles.div(nt) != n_idx);
      }
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      tc_fence_after();
      if (lane == 0) trace_ev(p, 1, tn, (tile << 8) | 0xff);   // accumulator stage free
      const uint32_t d_tmem = tmem_base + as * BN;
      if (PATCH) {
        bool first = true;
        for (int dx = 0; dx < 3; ++dx) {
          for (int cb = 0; cb < p.kb_per_tap; ++cb) {
            mbar_wait(&full_bar[pstage], pphase);
            for (int dy = 0; dy < 3; ++dy) {
              mbar_wait(&full_bar[PA + stage], phase);
              tc_fence_after();
              if (elect_one()) {
                // tap (dy, dx): the same patch, dy image rows further down (a whole number of swizzle atoms)
                const uint64_t adesc =
                    umma_desc_sw128_kmajor(smem_u32(sA + pstage * A_PATCH_BYTES + dy * p.patch_row_bytes));
                const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(sB + stage * Cfg::B_STAGE_BYTES));
#pragma unroll
                for (int k = 0; k < BLOCK_K / 16; ++k)
                  umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
                umma_commit(&empty_bar[PA + stage]);
                if (dy == 2) {
                  umma_commit(&empty_bar[pstage]);
                  if (dx == 2 && cb == p.kb_per_tap - 1) umma_commit(&tmem_full[as]);
                }
              }
              __syncwarp();
              first = false;
              if (++stage == PB) { stage = 0; phase ^= 1; }
            }
            if (++pstage == PA) { pstage = 0; pphase ^= 1; }
          }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
        continue;
      }
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) trace_ev(p, 1, tn, (tile << 8) | kb);     // operands landed, MMAs issued
        if (elect_one()) {
          const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(sA + stage * A_STAGE_BYTES));
          const uint64_t bdesc =
              umma_desc_sw128_kmajor(smem_u32(sB + (BSTAT ? kb : stage) * Cfg::B_STAGE_BYTES));
          if (CTA2) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k)
              umma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);              // the slot is free in both CTAs
            if (kb == kb_end - 1) umma_commit_pair(&tmem_full[as]);
          } else {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              // advance 16 fp16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in 16-byte units
              umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
            if (kb == kb_end - 1) {
              if (!RES) umma_commit(&tmem_full[as]);
              if (last_of_n) umma_commit(bres_empty);
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (RES) {
        constexpr uint32_t idesc16 = umma_idesc_f16(BLOCK_M, 16);
        const int nchunks = min(BN / 64, (p.cout - n_idx * BN + 63) / 64);
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t rdesc = umma_desc_sw128_kmajor(smem_u32(sA + stage * A_STAGE_BYTES));
            const uint64_t idesc_b = umma_desc_sw128_kmajor(smem_u32(sIdent));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d_tmem + c * 64 + k * 16, rdesc + 2 * k, idesc_b, idesc16, 1u);
            umma_commit(&empty_bar[stage]);
            if (c == nchunks - 1) umma_commit(&tmem_full[as]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else if (warp == 2 || warp == 3) {
    // ===================== TMA store warps =====================
    // One per epilogue group (warp 2 <-> group 0, warp 3 <-> group 1).  Each takes the staged 128x64 fp16 chunks of ITS
    // group (named barrier FULL<g>), issues the TMA store, waits until the store has read the staging buffer and hands
    // the buffer back (FREE<g>).  The two groups therefore never wait for each other: the first version had one store
    // warp that freed a buffer only after the OTHER group's next chunk had been issued, which serialised the groups
    // (device trace: 0.6 us per chunk spent waiting for the buffer, profiles/README.md).
    if (p.out_f32 == nullptr && !(p.dbg & 128)) {
      const int sg = warp - 2;
      int mine = 0, gb = 0;
      for (Walk wk = walk_begin(); walk_valid(wk); walk_next(wk)) {
        if (SK && wk.head) continue;              // parked head: finished (and stored) by the tail owner
        const int tile = wk.tile;
        int n_idx, m_idx, split;
        decode(tile, n_idx, m_idx, split);
        const int nchunks = min(BN / 64, (p.cout - n_idx * BN + 63) / 64);
        const int cf = (gb + sg) & 1;
        if (cf < nchunks) mine += (nchunks - cf + 1) / 2;
        gb += nchunks;
      }
      int done = 0;
      int stn = 0;
      gb = 0;
      for (Walk wk = walk_begin(); walk_valid(wk); walk_next(wk)) {
        if (SK && wk.head) continue;
        const int tile = wk.tile;
        int n_idx, m_idx, split;
        decode(tile, n_idx, m_idx, split);
        int tx, ty, img;
        locate(m_idx, tx, ty, img);
        const int x0 = tx * tw, y0 = ty * p.th;
        const int nchunks = min(BN / 64, (p.cout - n_idx * BN + 63) / 64);
        for (int c = (gb + sg) & 1; c < nchunks; c += 2) {
          named_bar_sync(BAR_FULL0 + sg, 160);
          // DVID_DBG=256 + DVID_TRACE: store warp 0 logs (chunk handed over / store issued / buffer read) in the load lane
          if ((p.dbg & 256) && sg == 0 && lane == 0) trace_ev(p, 0, stn, (tile << 8) | 0x10 | c);
          if (lane == 0 && !(p.dbg & 1)) {
            tma_store_4d(&tmC, sOut + sg * OUT_STAGE_BYTES, n_idx * BN + c * 64, x0, y0, img);
            tma_store_commit();
          }
          if ((p.dbg & 256) && sg == 0 && lane == 0) trace_ev(p, 0, stn, (tile << 8) | 0x20 | c);
          if (++done < mine) {                    // the group will stage another chunk into this buffer
            if (lane == 0) tma_store_wait_read<0>();
            if ((p.dbg & 256) && sg == 0 && lane == 0) trace_ev(p, 0, stn, (tile << 8) | 0x30 | c);
            __syncwarp();
            named_bar_arrive(BAR_FREE0 + sg, 160);
          }
        }
        gb += nchunks;
      }
      if (lane == 0) tma_store_wait<0>();
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // Global-memory latency is kept off the per-chunk critical path: the bias of the NEXT tile is fetched into
    // registers while the current tile is processed (and staged through smem), the residual of chunk c+1 is fetched
    // while chunk c is converted.
    // 8 warps = two groups of four; group g converts the 64-column chunks with (global chunk index & 1) == g into
    // staging buffer g.  Two warps per SM sub-partition hide each other's TMEM-load / smem / barrier latencies (the
    // 4-warp epilogue sat at ~0.15 IPC and, not the tensor pipe, bounded every K <= 512 layer).  The bias of the NEXT
    // tile is fetched while the current tile is processed; the residual of a group's next chunk while its current one
    // is converted.
    const int ew = (warp - 4) & 3;   // == warp % 4: the TMEM sub-partition this warp may read
    const int grp = (warp - 4) >> 2;
    const int row = ew * 32 + lane;
    const int et = threadIdx.x - 128;          // 0..255
    int gbase = 0;                             // 64-column chunks staged by both groups before this tile
    int as = 0;
    uint32_t aphase = 0;
    int tn = 0;
    const int gt = et & 127;                   // thread index inside the group
    int staged = 0;                            // chunks this group has handed to its store warp
    int staged_key = -1;                       // (n_idx, first chunk) whose bias is in the group's smem stage
    // bias of the group's chunks of a tile: thread gt holds column (gt & 63) of the group's (gt >> 6)-th chunk
    float bnext = 0.f;
    auto fetch_bias = [&](int tile, int gb) {
      int bn_idx, bm_idx, bsplit;
      decode(tile, bn_idx, bm_idx, bsplit);
      const int c = ((gb + grp) & 1) + 2 * (gt >> 6);
      const int col = bn_idx * BN + c * 64 + (gt & 63);
      bnext = (p.bias != nullptr && c * 64 < BN && col < p.cout) ? __ldg(p.bias + col) : 0.f;
    };
    if (!SK && tile_begin < tile_end && p.out_f32 == nullptr) fetch_bias(tile_begin, 0);
    for (Walk wk = walk_begin(); walk_valid(wk); walk_next(wk)) {
      const int tile = wk.tile;
      const bool sk_park = SK && wk.head;                 // (park) this CTA's trailing head segment, run first
      const bool sk_fin = SK && !wk.head && wk.kb0 > 0;  // (finish) leading tail segment: add the tile parked by CTA b-1
      int n_idx, m_idx, split;
      decode(tile, n_idx, m_idx, split);
      int tx, ty, img;
      locate(m_idx, tx, ty, img);
      const int x0 = tx * tw, y0 = ty * p.th;
      const int x = x0 + (row & (tw - 1));
      const int y = y0 + (row >> p.tw_log2);
      const bool valid = (x < p.w_out) && (y < p.h_out);
      const int nchunks = min(BN / 64, (p.cout - n_idx * BN + 63) / 64);
      const int c_first = (gbase + grp) & 1;   // first chunk of this tile that belongs to this group

      const __half* rrow = nullptr;   // this thread's residual row (channel 0)
      if (ERES && p.resid != nullptr && valid) {
        rrow = p.resid + ((static_cast<long long>(img) * p.resid_h + (y >> p.resid_shift)) * p.resid_w +
                          (x >> p.resid_shift)) * p.cout;
      }
      uint4 rnext[8];
      auto fetch_resid = [&](int c) {
        const int ch0 = n_idx * BN + c * 64;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          rnext[q] = (rrow != nullptr && ch0 + q * 8 < p.cout)
                         ? __ldg(reinterpret_cast<const uint4*>(rrow + ch0 + q * 8))
                         : make_uint4(0, 0, 0, 0);
        }
      };
      if (ERES && p.resid != nullptr && c_first < nchunks && !sk_park) fetch_resid(c_first);

      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      if (et == 0) trace_ev(p, 2, tn, (tile << 8) | 0xfe);       // accumulator ready
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;

      if (p.dbg & 128) {
        // experiment: no epilogue work at all (mainloop speed)
      } else if (sk_park) {
        // stream-K: park the fp32 accumulator tile; 32-column pieces alternate between the two groups, a warp writes
        // 32 rows x 16 B = 512 contiguous bytes per instruction
        float4* slot = reinterpret_cast<float4*>(p.sk_ws) + static_cast<size_t>(blockIdx.x) * (BN / 4) * BLOCK_M;
#pragma unroll 1
        for (int c = grp; c < BN / 32; c += 2) {
          uint32_t v[32];
          tmem_ld32(tbase + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            slot[(c * 8 + q) * BLOCK_M + row] =
                make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                            __uint_as_float(v[4 * q + 3]));
          }
        }
        __threadfence();
        named_bar_sync(BAR_SK, 256);
        if (et == 0) st_release_gpu(p.sk_flags + blockIdx.x, 1u);
        if (et == 0) trace_ev(p, 2, tn, (tile << 8) | 0xfd);     // partial tile parked
      } else if (p.out_f32 != nullptr) {
        // split-K partials: fp32, direct vector stores (GEMM view: th == 1, row index = x); 32-column pieces
        // alternate between the two groups
        float* dst = p.out_f32 + (static_cast<long long>(split) * p.m_total + x) * p.cout;
#pragma unroll 1
        for (int c = grp; c < BN / 32; c += 2) {
          const int ch0 = n_idx * BN + c * 32;
          if (ch0 >= p.cout) break;
          uint32_t v[32];
          tmem_ld32(tbase + c * 32, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (ch0 + q * 4 < p.cout) {
                float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                       __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                *reinterpret_cast<float4*>(dst + ch0 + q * 4) = o;
              }
            }
          }
        }
      } else {
        // stage this tile's bias: the first barrier orders the previous tile's last sBias reads before these writes,
        // the second these writes before this tile's reads.  Then start fetching the next tile's bias.
        float* gBias = sBias + grp * (BN > 128 ? BN / 2 : 64);   // each group stages its own chunks' bias: no barrier
        const int bias_key = n_idx * 2 + c_first;                // across the groups
        if (SK && bias_key != staged_key) fetch_bias(tile, gbase);   // stream-K: no look-ahead (one key per CTA)
        if (bias_key != staged_key) {      // the weight-stationary walk keeps n_idx for many tiles: stage once
          named_bar_sync(BAR_BIAS0 + grp, 128);
          if (gt < (BN > 128 ? BN / 2 : 64)) gBias[gt] = bnext;
          named_bar_sync(BAR_BIAS0 + grp, 128);
          staged_key = bias_key;
        }
        if (!SK && tile + tile_step < tile_end) fetch_bias(tile + tile_step, gbase + nchunks);
        uint8_t* buf = sOut + grp * OUT_STAGE_BYTES;
        const float4* parked = nullptr;
        if (sk_fin) {       // CTA b-1 parked the head part of this tile as its first item
          const unsigned* flag = p.sk_flags + blockIdx.x - 1;
          while (ld_acquire_gpu(flag) == 0u) {}
          if (et == 0) trace_ev(p, 2, tn, (tile << 8) | 0xfc);   // parked tile of CTA b-1 visible
          parked = reinterpret_cast<const float4*>(p.sk_ws) + static_cast<size_t>(blockIdx.x - 1) * (BN / 4) * BLOCK_M;
        }
        // the accumulator columns of the group's NEXT chunk are requested as soon as the current chunk is converted, so
        // the TMEM load latency runs under the staging-buffer wait, the shared-memory stores and the hand-over
        uint32_t v[2][32];
        if (c_first < nchunks) {
          tmem_ld32(tbase + c_first * 64, v[0]);
          tmem_ld32(tbase + c_first * 64 + 32, v[1]);
        }
#pragma unroll 1
        for (int c = c_first; c < nchunks; c += 2) {
          uint4 rcur[8];
          if (ERES && p.resid != nullptr) {
#pragma unroll
            for (int q = 0; q < 8; ++q) rcur[q] = rnext[q];
            if (c + 2 < nchunks) fetch_resid(c + 2);
          }
          float4 pk[SK ? 16 : 1];
          if (SK && sk_fin) {                    // all 16 loads of the chunk in flight (L2 hits, written by another SM)
#pragma unroll
            for (int q = 0; q < 16; ++q) pk[q] = __ldcg(parked + (c * 16 + q) * BLOCK_M + row);
          }
          tmem_ld_wait();
          if (SK && sk_fin) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              uint32_t* vv = &v[q >> 3][(q & 7) * 4];
              vv[0] = __float_as_uint(__uint_as_float(vv[0]) + pk[q].x);
              vv[1] = __float_as_uint(__uint_as_float(vv[1]) + pk[q].y);
              vv[2] = __float_as_uint(__uint_as_float(vv[2]) + pk[q].z);
              vv[3] = __float_as_uint(__uint_as_float(vv[3]) + pk[q].w);
            }
          }
          // The whole chunk is converted into registers BEFORE waiting for the staging buffer: the store warp hands the
          // buffer back only after the TMA store has read it (~0.6 us behind the hand-over in the device trace, the TMA unit
          // is busy with the operand loads), and that wait used to sit in front of the arithmetic.
          uint4 packed[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(gBias + ((c - c_first) >> 1) * 64 + h * 32 + j);
              f[j] = __uint_as_float(v[h][j]) + b4.x;
              f[j + 1] = __uint_as_float(v[h][j + 1]) + b4.y;
              f[j + 2] = __uint_as_float(v[h][j + 2]) + b4.z;
              f[j + 3] = __uint_as_float(v[h][j + 3]) + b4.w;
            }
            if (ERES && p.resid != nullptr) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const __half2* rh = reinterpret_cast<const __half2*>(&rcur[h * 4 + q]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 rf = __half22float2(rh[e]);
                  f[q * 8 + 2 * e] += rf.x;
                  f[q * 8 + 2 * e + 1] += rf.y;
                }
              }
            }
            if (p.relu == 2) {   // GELU (erf form, torch.nn.GELU default) - Swin MLP, swintransformer.py:47-66
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
            }
            const __half2 zero2 = __float2half2_rn(0.f);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              __half2 h2[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                h2[e] = __floats2half2_rn(f[q * 8 + 2 * e], f[q * 8 + 2 * e + 1]);
                if (p.relu == 1) h2[e] = __hmax2(h2[e], zero2);   // ReLU commutes with the rounding to fp16
              }
              packed[h * 4 + q] = *reinterpret_cast<uint4*>(h2);
            }
          }
          if (c + 2 < nchunks) {
            tmem_ld32(tbase + (c + 2) * 64, v[0]);
            tmem_ld32(tbase + (c + 2) * 64 + 32, v[1]);
          }
          // the store warp has drained the TMA store that last read this group's staging buffer
          if (staged > 0) named_bar_sync(BAR_FREE0 + grp, 160);
#pragma unroll
          for (int chunk = 0; chunk < 8; ++chunk)   // 16-byte chunk inside the 128-byte row
            *reinterpret_cast<uint4*>(buf + row * 128 + ((chunk ^ (row & 7)) << 4)) = packed[chunk];
          fence_proxy_async_smem();
          named_bar_arrive(BAR_FULL0 + grp, 160);   // hand the staged chunk to the store warp, do not wait for it
          ++staged;
          if (et == 0) trace_ev(p, 2, tn, (tile << 8) | c);
        }
        gbase += nchunks;
        if (sk_fin) {       // every thread has consumed its part of the parked tile: re-arm the flag for the next launch
          named_bar_sync(BAR_SK, 256);
          if (et == 0) st_release_gpu(p.sk_flags + blockIdx.x - 1, 0u);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2) mbar_arrive_leader(&tmem_empty[as]);     // the leader's MMA warp waits for both CTAs' epilogues
        else mbar_arrive(&tmem_empty[as]);
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  if (CTA2) cluster_sync_all();      // nothing of the pair is in flight towards either CTA any more
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ host side

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
      return nullptr;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// fp16 tensor map, 128-byte swizzle, dims innermost-first. Returns 0 on success.
int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, const uint32_t* elem_strides) {
  PFN_encodeTiled enc = get_encode();
  if (enc == nullptr) return DVID_ERR_DRIVER;
  cuuint64_t gdim[5];
  cuuint64_t gstride[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstride[i] = strides_bytes[i];
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstride, bdim, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "dvid: cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,...] box=[%u,%u,...]\n", (int)r,
            rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
            rank > 1 ? box[1] : 0);
    return DVID_ERR_DRIVER;
  }
  return 0;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DVID_PDL");
    v = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return v != 0;
}

static int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// stream-K workspace: one fp32 accumulator tile (128 x 256) + one flag per CTA.  Allocated once, outside any stream
// capture, by conv_streamk_enable(); a single workspace serves every launch because the launches of one stream are
// ordered (the kernel touches it only after griddepcontrol.wait) - callers that run convolutions on several streams at
// once must switch it off.
static float* g_sk_ws = nullptr;
static unsigned* g_sk_flags = nullptr;
static bool g_sk_on = false;

int conv_streamk_enable(int on) {
  if (on && g_sk_ws == nullptr) {
    const size_t ws_bytes = static_cast<size_t>(num_sms()) * BLOCK_M * 256 * sizeof(float);
    const size_t flag_bytes = static_cast<size_t>(num_sms() + 1) * sizeof(unsigned);
    void* ptr = nullptr;
    if (cudaMalloc(&ptr, ws_bytes + flag_bytes) != cudaSuccess) return DVID_ERR_CUDA;
    g_sk_ws = static_cast<float*>(ptr);
    g_sk_flags = reinterpret_cast<unsigned*>(static_cast<uint8_t*>(ptr) + ws_bytes);
    if (cudaMemset(g_sk_flags, 0, flag_bytes) != cudaSuccess) return DVID_ERR_CUDA;
  }
  g_sk_on = on != 0;
  return 0;
}

// CTA-pair variant: {2,1,1} clusters, one pair per TPC, the pairs walk pairs of M tiles
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const ConvGemmParams& p,
                       cudaStream_t stream) {
  using Cfg = ConvGemmCfg<256, false, false, false, true>;
  auto kernel = conv_gemm_kernel<256, false, false, false, false, true>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess)
      return DVID_ERR_CUDA;
    attr_set = true;
  }
  const int pair_tiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
  const int pairs = pair_tiles < num_sms() / 2 ? pair_tiles : num_sms() / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmC, tmC, p);
  return cudaGetLastError() == cudaSuccess ? 0 : DVID_ERR_CUDA;
}

template <int BN, bool BSTAT, bool RES, bool SK = false, bool PATCH = false>
static int launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                      const ConvGemmParams& p, cudaStream_t stream) {
  using Cfg = ConvGemmCfg<BN, BSTAT, RES, PATCH>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BN, BSTAT, RES, SK, PATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return DVID_ERR_CUDA;
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.splits;
  const int grid = total < num_sms() ? total : num_sms();
  if (getenv("DVID_TRACE") != nullptr) {     // debug: run once with the event log and print it (synchronises!)
    ConvGemmParams q = p;
    const size_t bytes = 3 * 2048 * sizeof(unsigned long long);
    cudaMalloc(&q.trace, bytes);
    cudaMemsetAsync(q.trace, 0, bytes, stream);
    launch_pdl(conv_gemm_kernel<BN, BSTAT, RES, SK, PATCH>, dim3(grid), dim3(384), Cfg::SMEM_BYTES, stream, tmA, tmB, tmC, tmR, q);
    cudaStreamSynchronize(stream);
    static unsigned long long host[3 * 2048];
    cudaMemcpy(host, q.trace, bytes, cudaMemcpyDeviceToHost);
    cudaFree(q.trace);
    unsigned long long t0 = ~0ull;
    for (int r = 0; r < 3; ++r)
      if (host[r * 2048 + 1] && host[r * 2048 + 1] < t0) t0 = host[r * 2048 + 1];
    fprintf(stderr, "dvid trace BN=%d bstat=%d tiles=%d grid=%d total_kb=%d\n", BN, (int)BSTAT, total, grid, p.total_kb);
    const char* names[3] = {"load", "mma ", "epi "};
    for (int r = 0; r < 3; ++r)
      for (int i = 0; i < 1023 && host[r * 2048 + 2 * i + 1]; ++i) {
        if (i > 120) break;
        fprintf(stderr, "  %s tile %4llu ev %3llu  t=%8.2f us\n", names[r], host[r * 2048 + 2 * i] >> 8,
                host[r * 2048 + 2 * i] & 0xff, (host[r * 2048 + 2 * i + 1] - t0) / 1900.0);
      }
    return cudaGetLastError() == cudaSuccess ? 0 : DVID_ERR_CUDA;
  }
  launch_pdl(conv_gemm_kernel<BN, BSTAT, RES, SK, PATCH>, dim3(grid), dim3(384), Cfg::SMEM_BYTES, stream, tmA, tmB, tmC, tmR, p);
  return cudaGetLastError() == cudaSuccess ? 0 : DVID_ERR_CUDA;
}

// in: NHWC fp16 [n,h,w,cin]; weight: [cout][R*S*cin] fp16; out: NHWC fp16 [n,h_out,w_out,cout] (or out_f32 partials).
int conv_gemm_launch(const void* in, const void* weight, const float* bias, const void* resid, void* out,
                     float* out_f32, int n, int h, int w, int cin, int cout, int R, int S, int stride, int pad,
                     int resid_shift, int relu, int splits, int force_bn, cudaStream_t stream,
                     const uint64_t* a_strides_bytes) {
  if (cin % 8 != 0 || cout % 8 != 0) return DVID_ERR_SHAPE;
  if (n <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0) return DVID_ERR_SHAPE;
  if (stride < 1 || stride > 2) return DVID_ERR_SHAPE;
  ConvGemmParams p;
  p.n_img = n;
  p.h_out = (h + 2 * pad - R) / stride + 1;
  p.w_out = (w + 2 * pad - S) / stride + 1;
  if (p.h_out <= 0 || p.w_out <= 0) return DVID_ERR_SHAPE;
  p.cin = cin;
  p.cout = cout;
  p.R = R; p.S = S; p.pad = pad; p.stride = stride;
  // choose the th x tw pixel patch (tw power of two, th*tw = 128) that covers the image with fewest tiles
  int best_log2 = 7;
  long long best_tiles = -1;
  for (int l2 = 7; l2 >= 0; --l2) {
    const int tw = 1 << l2, th = 128 >> l2;
    if (stride > 1 && (tw * stride > 256 || th * stride > 256)) continue;
    const long long t = static_cast<long long>((p.w_out + tw - 1) / tw) * ((p.h_out + th - 1) / th);
    if (best_tiles < 0 || t < best_tiles) { best_tiles = t; best_log2 = l2; }
  }
  if (best_tiles < 0) return DVID_ERR_SHAPE;
  // column-patch walk (see ConvGemmCfg): plain 3x3 / stride 1 / pad 1 convolutions.  The tile must be >= 4 rows tall for
  // the shared patch to pay and >= 8 pixels wide for the row views to stay swizzle-atom aligned; fewest tiles first, the
  // taller tile on ties (less A traffic).  The choice depends on the image size only (batch invariance).
  static int patch_env = -1, cta2_env = -1;
  if (patch_env < 0) { const char* e = getenv("DVID_CONV_PATCH"); patch_env = e ? atoi(e) : 0; }
  if (cta2_env < 0) { const char* e = getenv("DVID_CONV_CTA2"); cta2_env = e ? atoi(e) : 1; }
  // (with CTA pairs enabled the 256-wide layers go to them: the patch walk has no deeper ring to offer at that width)
  const bool patch = patch_env && R == 3 && S == 3 && stride == 1 && pad == 1 && out != nullptr && resid == nullptr &&
                     a_strides_bytes == nullptr && (!cta2_env || cout < 256);
  if (patch) {
    best_tiles = -1;
    for (int l2 = 5; l2 >= 3; --l2) {                 // (th, tw) = (4, 32), (8, 16), (16, 8)
      const int tw = 1 << l2, th = 128 >> l2;
      const long long t = static_cast<long long>((p.w_out + tw - 1) / tw) * ((p.h_out + th - 1) / th);
      if (best_tiles < 0 || t <= best_tiles) { best_tiles = t; best_log2 = l2; }
    }
  }
  p.tw_log2 = best_log2;
  p.th = 128 >> best_log2;
  const int tw = 1 << best_log2;
  p.tiles_x = (p.w_out + tw - 1) / tw;
  p.tiles_y = (p.h_out + p.th - 1) / p.th;
  p.m_tiles = n * p.tiles_x * p.tiles_y;
  p.kb_per_tap = (cin + BLOCK_K - 1) / BLOCK_K;
  p.total_kb = R * S * p.kb_per_tap;
  if (out_f32 == nullptr) splits = 1;
  if (splits < 1) splits = 1;
  if (splits > p.total_kb) splits = p.total_kb;
  p.kb_per_split = (p.total_kb + splits - 1) / splits;
  p.splits = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
  p.bias = bias;
  p.resid = reinterpret_cast<const __half*>(resid);
  p.resid_shift = resid_shift;
  p.resid_h = ((p.h_out - 1) >> resid_shift) + 1;
  p.resid_w = ((p.w_out - 1) >> resid_shift) + 1;
  p.relu = relu;
  p.out_f32 = out_f32;
  p.m_total = static_cast<long long>(n) * p.h_out * p.w_out;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("DVID_DBG"); dbg = e ? atoi(e) : 0; }
    p.dbg = dbg;
    p.trace = nullptr;
    { const char* e = getenv("DVID_TRACE_CTA"); p.trace_cta = e ? atoi(e) : 0; }
  }
  if (out_f32 != nullptr && !(p.h_out == 1 && n == 1 && p.th == 1)) return DVID_ERR_SHAPE;  // GEMM view only

  int bn = force_bn;
  {
    static int env_bn = -1;
    if (env_bn < 0) { const char* e = getenv("DVID_FORCE_BN"); env_bn = e ? atoi(e) : 0; }
    if (bn == 0 && env_bn > 0 && cout >= env_bn) bn = env_bn;
  }
  // Batch invariance (a frame's result must not depend on how many other frames share the launch): the
  // tile WIDTH may follow the tile count - every output element is accumulated over K in the same k-block order for any
  // width - but the way a same-resolution residual is added may not: >= 128-wide tiles add it through the tensor core
  // (identity MMA into the accumulator, before the bias), 64-wide tiles in the epilogue registers (after the bias), and
  // the two round differently.  So convolutions with such a residual never take the 64-wide tile.
  const bool res_same = resid != nullptr && resid_shift == 0 && out != nullptr && cout >= 128;
  if (res_same && bn == 64) bn = 128;
  static int cost_model = -1;
  if (cost_model < 0) { const char* e = getenv("DVID_BN_COST"); cost_model = e ? atoi(e) : 1; }
  if (bn == 0 && cost_model) {
    // Tile width by a wave-quantisation cost model: waves(bn) x relative tile cost.  The costs are measured on B200
    // (tools/bench_gemm.py with DVID_FORCE_BN): a 128-wide tile costs ~0.78 of a 256-wide one (the A tile is re-read
    // from smem per MMA, same epilogue overheads), a 64-wide one ~0.55.  Example: res5 3x3 conv, 38 M tiles x 512
    // channels: bn=256 -> 1 wave x 1.0 (37.9 us), bn=128 -> 2 waves x 0.78 (52.2 us).
    // (Round-2 device traces explain the costs: an M=128 x K=16 MMA occupies the tensor pipe for ~130 cycles whatever its
    // N, so a k-block of a 64-wide tile takes as long as one of a 256-wide tile - 0.27-0.3 us - and only the epilogue and
    // the bytes per k-block shrink with the width.)
    const double tile_cost[3] = {1.0, 0.78, 0.55};
    const int cand[3] = {256, 128, 64};
    double best = 1e30;
    for (int i = 0; i < 3; ++i) {
      if (cand[i] > 64 && cout < cand[i]) continue;
      if (res_same && cand[i] < 128) continue;
      const long long tiles = static_cast<long long>(p.m_tiles) * p.splits * ((cout + cand[i] - 1) / cand[i]);
      const long long waves = (tiles + num_sms() - 1) / num_sms();
      const double c = static_cast<double>(waves) * tile_cost[i];
      if (c < best * 0.97) { best = c; bn = cand[i]; }     // prefer the wider tile unless clearly worse
    }
  }
  if (bn == 0) {
    const long long want = (num_sms() * 4) / 5;
    if (cout >= 256 && static_cast<long long>(p.m_tiles) * p.splits * ((cout + 255) / 256) >= want) bn = 256;
    else if (cout >= 128 && static_cast<long long>(p.m_tiles) * p.splits * ((cout + 127) / 128) >= want) bn = 128;
    else bn = (cout >= 256 && p.m_tiles * p.splits >= 16) ? 128 : 64;
    if (cout <= 64) bn = 64;
    if (res_same && bn == 64) bn = 128;
  }
  if (bn != 64 && bn != 128 && bn != 256) return DVID_ERR_SHAPE;
  if (bn == 256 && resid != nullptr && !(resid_shift == 0 && out != nullptr)) bn = 128;   // epilogue residual: see ERES
  // K <= 256 with a tensor-core residual (bottleneck conv3 of res2..res4): the four residual chunks of a 256-wide tile
  // queue behind its four k-blocks in a THREE-stage ring (128 KB of resident weights leave no room for more) and the
  // device trace shows 2.2 us per tile between the last k-block and the finished accumulator; the 128-wide tile has a
  // five-stage ring: 26.4 -> 24.6 us (res4), 39.4 -> 36.0 (res3), 64.8 -> 60.0 (res2), tools/bench_conv_res.py
  if (bn == 256 && res_same && p.total_kb <= BSTAT_MAX_KB && force_bn == 0) {
    static int env_forced = -1;
    if (env_forced < 0) { const char* e = getenv("DVID_FORCE_BN"); env_forced = (e && atoi(e) > 0) ? 1 : 0; }
    if (!env_forced) bn = 128;
  }
  p.n_tiles = (cout + bn - 1) / bn;
  p.div_m_tiles.init(p.m_tiles);
  p.div_n_tiles.init(p.n_tiles);
  p.div_tiles_x.init(p.tiles_x);
  p.div_tiles_y.init(p.tiles_y);
  p.div_total_kb.init(p.total_kb);
  {
    static int wpf = -1;
    if (wpf < 0) { const char* e = getenv("DVID_WPREFETCH"); wpf = e ? atoi(e) : 1; }
    p.w_base = static_cast<const uint8_t*>(weight);
    p.w_bytes = static_cast<unsigned long long>(cout) * R * S * cin * 2;
    p.w_slice_bytes = 0;
    { static int bpre = -1; if (bpre < 0) { const char* e = getenv("DVID_BPRE"); bpre = e ? atoi(e) : 1; } p.b_pre = bpre; }
    if (wpf && p.w_bytes <= (8ull << 20) && (reinterpret_cast<uintptr_t>(weight) & 15) == 0) {
      const long long total = static_cast<long long>(p.m_tiles) * p.n_tiles * p.splits;
      const unsigned long long ctas = static_cast<unsigned long long>(total < num_sms() ? total : num_sms());
      const unsigned long long per = (p.w_bytes + ctas - 1) / ctas;
      p.w_slice_bytes = static_cast<unsigned>((per + 127) & ~127ull);
    }
  }
  p.sk_ws = nullptr;
  p.sk_flags = nullptr;
  p.patch_row_bytes = tw * 128;
  p.patch_bytes = (p.th + 2) * tw * 128;

  // CTA pairs (ConvGemmCfg): plain 256-wide convolutions with K deep enough for the ring to matter
  // (measured per layer, tools/bench_conv_patch.py: -5..-9 % from 16 k-blocks up, +5 % at 8)
  const bool cta2 = cta2_env && !patch && bn == 256 && out != nullptr && resid == nullptr && p.splits == 1 &&
                    p.total_kb >= 16 && p.m_tiles >= 2 && p.trace == nullptr && p.dbg == 0;
  CUtensorMap tmA, tmB, tmC;
  {
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    uint64_t strides[3] = {(uint64_t)cin * 2, (uint64_t)w * cin * 2, (uint64_t)h * w * cin * 2};
    if (a_strides_bytes != nullptr) {   // overlapping-window view (stem convolution), see stem_conv_launch
      for (int i = 0; i < 3; ++i) strides[i] = a_strides_bytes[i];
    }
    const uint32_t box[4] = {(uint32_t)BLOCK_K, (uint32_t)(tw * stride), (uint32_t)((patch ? p.th + 2 : p.th) * stride), 1};
    const uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    int r = make_tmap_f16(&tmA, in, 4, dims, strides, box, es);
    if (r) return r;
  }
  {
    const uint64_t ktot = (uint64_t)R * S * cin;
    const uint64_t dims[2] = {ktot, (uint64_t)cout};
    const uint64_t strides[1] = {ktot * 2};
    const uint32_t box[2] = {(uint32_t)BLOCK_K, (uint32_t)(cta2 ? bn / 2 : bn)};
    int r = make_tmap_f16(&tmB, weight, 2, dims, strides, box, nullptr);
    if (r) return r;
  }
  if (out != nullptr) {
    const uint64_t dims[4] = {(uint64_t)cout, (uint64_t)p.w_out, (uint64_t)p.h_out, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)cout * 2, (uint64_t)p.w_out * cout * 2,
                                 (uint64_t)p.h_out * p.w_out * cout * 2};
    const uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)p.th, 1};
    int r = make_tmap_f16(&tmC, out, 4, dims, strides, box, nullptr);
    if (r) return r;
  } else {
    if (out_f32 == nullptr) return DVID_ERR_SHAPE;
    tmC = tmA;  // unused by the fp32 path
  }
  if (cta2) return launch_pair(tmA, tmB, tmC, p, stream);
  if (patch) {
    if (p.splits != 1 || p.patch_bytes > A_PATCH_BYTES) return DVID_ERR_SHAPE;
    if (bn == 256) return launch_cfg<256, false, false, false, true>(tmA, tmB, tmC, tmC, p, stream);
    if (bn == 128) return launch_cfg<128, false, false, false, true>(tmA, tmB, tmC, tmC, p, stream);
    return launch_cfg<64, false, false, false, true>(tmA, tmB, tmC, tmC, p, stream);
  }
  // weight-stationary walk when the whole K fits (<= 256) and every CTA gets several tiles of the same weight tile
  static int bstat_env = -1;
  if (bstat_env < 0) { const char* e = getenv("DVID_BSTAT"); bstat_env = e ? atoi(e) : 1; }
  const bool bstat = bstat_env && out_f32 == nullptr && p.total_kb <= BSTAT_MAX_KB && bn >= 128 &&
                     static_cast<long long>(p.m_tiles) * p.n_tiles >= 2LL * num_sms();
  // residual through the tensor core (identity MMA) whenever it is a same-resolution tensor and the tile is >= 128 wide
  const bool res_mma = resid != nullptr && resid_shift == 0 && out != nullptr && bn >= 128;
  CUtensorMap tmR = tmC;
  if (res_mma) {
    const uint64_t dims[4] = {(uint64_t)cout, (uint64_t)p.w_out, (uint64_t)p.h_out, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)cout * 2, (uint64_t)p.w_out * cout * 2,
                                 (uint64_t)p.h_out * p.w_out * cout * 2};
    const uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)p.th, 1};
    int r = make_tmap_f16(&tmR, resid, 4, dims, strides, box, nullptr);
    if (r) return r;
  }
  if (bstat) {
    if (res_mma) {
      if (bn == 256) return launch_cfg<256, true, true>(tmA, tmB, tmC, tmR, p, stream);
      return launch_cfg<128, true, true>(tmA, tmB, tmC, tmR, p, stream);
    }
    if (bn == 256) return launch_cfg<256, true, false>(tmA, tmB, tmC, tmR, p, stream);
    return launch_cfg<128, true, false>(tmA, tmB, tmC, tmR, p, stream);
  }
  if (res_mma) {
    if (bn == 256) return launch_cfg<256, false, true>(tmA, tmB, tmC, tmR, p, stream);
    return launch_cfg<128, false, true>(tmA, tmB, tmC, tmR, p, stream);
  }
  {
    // stream-K when whole-tile scheduling would waste >= 10 % of the launch (idle part of the last wave / waves): 152
    // tiles -> 49 %, 608 -> 18 %, 2432 -> 3 % (not worth the parked-tile round trip).  tiles >= SMs keeps every CTA's
    // range at least one tile long (a tile is shared by two CTAs at most); the flattened index must fit an int.
    const long long tiles = static_cast<long long>(p.m_tiles) * p.n_tiles;
    const int sms = num_sms();
    const long long waves = (tiles + sms - 1) / sms;
    const double idle = static_cast<double>(waves * sms - tiles) / sms;
    static int sk_min_kb = -1;
    static int sk_max_waves = -1;
    if (sk_min_kb < 0) { const char* e = getenv("DVID_SK_MIN_KB"); sk_min_kb = e ? atoi(e) : 32; }
    if (sk_max_waves < 0) { const char* e = getenv("DVID_SK_MAX_WAVES"); sk_max_waves = e ? atoi(e) : 2; }
    if (g_sk_on && g_sk_ws != nullptr && out != nullptr && p.splits == 1 && tiles >= sms && idle >= 0.1 * waves &&
        waves <= sk_max_waves &&
        p.total_kb >= sk_min_kb && tiles * p.total_kb < (1LL << 30) && p.trace == nullptr && p.dbg == 0) {
      p.sk_ws = g_sk_ws;
      p.sk_flags = g_sk_flags;
      if (bn == 256) return launch_cfg<256, false, false, true>(tmA, tmB, tmC, tmR, p, stream);
      if (bn == 128) return launch_cfg<128, false, false, true>(tmA, tmB, tmC, tmR, p, stream);
      return launch_cfg<64, false, false, true>(tmA, tmB, tmC, tmR, p, stream);
    }
  }
  if (bn == 256) return launch_cfg<256, false, false>(tmA, tmB, tmC, tmR, p, stream);
  if (bn == 128) return launch_cfg<128, false, false>(tmA, tmB, tmC, tmR, p, stream);
  return launch_cfg<64, false, false>(tmA, tmB, tmC, tmR, p, stream);
}

// Stem convolution 7x7 / stride 2 / pad 3 with 3 input channels (detectron2 BasicStem) as an implicit
// GEMM without im2col.  Input: the zero-haloed NHWC8 image written by preprocess_launch, [n][H+6][W+6][8] fp16
// (3 real channels).  A TMA view with OVERLAPPING rows exposes, for every pixel x of a padded row, the 64 contiguous
// fp16 = 8 pixels x 8 channels starting there: dims {64, W-1, H+6, n}, strides {16 B, (W+6)*16 B, ...}.  One filter
// row (r) is then a K=64 block, so the 7x7x3 filter becomes a 7x1 "convolution" over a 64-channel virtual image
// with the weights laid out [cout][r][s(8, last zero)][c(8, last five zero)].
int stem_conv_launch(const void* in_haloed, const void* weight, const float* bias, void* out, int n, int H, int W,
                     int cout, int relu, cudaStream_t stream) {
  if (H % 2 != 0 || W % 2 != 0 || H < 8 || W < 8) return DVID_ERR_SHAPE;
  const uint64_t wp = (uint64_t)W + 6, hp = (uint64_t)H + 6;
  const uint64_t strides[3] = {16, wp * 16, hp * wp * 16};
  return conv_gemm_launch(in_haloed, weight, bias, nullptr, out, nullptr, n, H + 6, W - 1, 64, cout, 7, 1, 2, 0, 0, relu,
                          1, 0, stream, strides);
}

}  // namespace dvid
