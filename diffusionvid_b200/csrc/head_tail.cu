// Fused classification / regression towers + predictors + apply_deltas of one RCNNHead evaluation.
//
// Reference: mega_core/modeling/roi_heads/box_head/box_head.py:538-590 (RCNNHead.forward tail, identical in
// RCNNHead_cond :649-664):
//     cls = ReLU(LN(fc @ Wc^T))                      cls_module (NUM_CLS = 1)
//     logits = cls @ Wl^T + bl                       class_logits
//     reg = fc; 3 x reg = ReLU(LN(reg @ Wr_i^T))     reg_module (NUM_REG = 3)
//     deltas = reg @ Wd^T + bd ; boxes = apply_deltas(deltas, boxes)
// The reference runs this as 6 cuBLAS GEMMs + 4 LayerNorm + 4 ReLU + ~15 elementwise kernels per head; the first
// version here used 6 tcgen05 GEMM launches + 5 row kernels (~100 us at M = 2400, all launch/latency bound: each GEMM
// is 0.3 GFLOP).  This kernel keeps a 128-row tile of activations in shared memory for the whole chain: one CTA per
// 128 boxes, the six weight matrices stream through a TMA ring, tcgen05.mma accumulates in TMEM, the epilogue warps
// do LayerNorm + ReLU straight out of TMEM and write the next layer's A operand back to shared memory in the
// 128-byte-swizzled K-major layout the MMA descriptors expect.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..11 = epilogue: two threads per row (a row is a
// TMEM lane; warp w reads sub-partition w % 4, warps 4..7 own columns 0..127 and warps 8..11 columns 128..255 of it).
// The six layers form a strict chain, so the kernel's time is 6 x (weight wait + MMA + epilogue): the first version's
// epilogue (one thread per row, three passes of eight serialised 32-column TMEM loads) cost ~6 us of each ~8 us LN
// layer; now each thread makes two passes over its 128 columns, two 32-column loads in flight, statistics in one pass
// (sum and sum of squares) exchanged between the two column halves through shared memory.
#include "ptx_sm100.cuh"
#include "dvid_internal.h"

namespace dvid {

namespace {

constexpr int TM = 128;                 // rows per CTA
constexpr int D = 256;                  // hidden width
constexpr int KB = D / 64;              // k-blocks of 64 per layer
constexpr int A_KB_BYTES = TM * 128;    // 16 KB: one k-block of an activation tile
constexpr int A_BUF_BYTES = KB * A_KB_BYTES;   // 64 KB
constexpr int W_STAGE_BYTES = 256 * 128;       // 32 KB: one k-block of a 256-row weight matrix
constexpr int W_STAGES = 2;
constexpr int NLAYERS = 6;
constexpr int NL_PAD = 32;              // class_logits rows padded to 32
constexpr int ND_PAD = 16;              // bboxes_delta rows padded to 16
constexpr int COL_ACC = 0, COL_LOGIT = 256, COL_DELTA = 288;
constexpr int EPI_THREADS = 256;        // warps 4..11
constexpr int SMEM_BYTES = 2 * A_BUF_BYTES + W_STAGES * W_STAGE_BYTES + 2 * D * 4 /*gamma,beta*/ + 256 +
                           2 * TM * 8 /*row statistics of the two column halves*/ + 1024;
constexpr float kScaleClamp = 8.740336742730447f;   // log(100000/16), box_head.py _DEFAULT_SCALE_CLAMP

struct TailParams {
  int M, C;
  const float* ln_g[4];
  const float* ln_b[4];
  const float* cls_bias;
  const float* delta_bias;
  const float* boxes_in;
  float* logits_out;
  float* boxes_out;
  const uint8_t* w_ptr[NLAYERS];   // the six weight matrices ([n_l][256] fp16) for the pre-wait L2 prefetch
};

struct TailMaps {
  CUtensorMap a;
  CUtensorMap w[NLAYERS];
};

__device__ __forceinline__ int layer_n(int l) { return l == 1 ? NL_PAD : (l == 5 ? ND_PAD : D); }

__global__ void __launch_bounds__(384, 1)
head_tail_kernel(const __grid_constant__ TailMaps tm, const TailParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sAct = smem;                                  // 2 x 64 KB activation tiles (A operands)
  uint8_t* sW = sAct + 2 * A_BUF_BYTES;                  // weight ring
  float* sG = reinterpret_cast<float*>(sW + W_STAGES * W_STAGE_BYTES);
  float* sBt = sG + D;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(sBt + D);
  uint64_t* w_empty = w_full + W_STAGES;
  uint64_t* a_full = w_empty + W_STAGES;
  uint64_t* acc_full = a_full + 1;
  uint64_t* epi_done = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_done + 1);
  float2* sStat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(w_full) + 256);   // [2][TM] (sum, sum of squares)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TM;
  pdl_trigger();

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm.a);
    for (int l = 0; l < NLAYERS; ++l) tma_prefetch_desc(&tm.w[l]);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < W_STAGES; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(acc_full, 1);
    mbar_init(epi_done, EPI_THREADS);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  if (warp == 3 && lane < NLAYERS) {
    // The weights do not depend on the previous kernel: ask L2 for them before the dependency wait.  Between two
    // evaluations of the same head ~1.5 GB stream through L2, so they come from HBM, and with a 2-stage ring the second
    // half of every layer's weights would otherwise be requested only when the first half has been consumed
    // (~2 us of exposed HBM latency per layer, 6 layers).  Each CTA prefetches its 1/gridDim.x slice of every matrix.
    const unsigned bytes = static_cast<unsigned>(layer_n(lane)) * D * 2;
    const unsigned slice = ((bytes + gridDim.x - 1) / gridDim.x + 127u) & ~127u;
    const unsigned off = blockIdx.x * slice;
    if (off < bytes) {
      const unsigned n = min(slice, bytes - off) & ~15u;
      if (n != 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.w_ptr[lane] + off), "r"(n) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      // ===================== TMA producer: the activation tile, then 6 x 4 weight k-blocks =====================
      mbar_expect_tx(a_full, A_BUF_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sAct + kb * A_KB_BYTES, &tm.a, a_full, kb * 64, m0);
      int stage = 0;
      uint32_t phase = 0;
      for (int l = 0; l < NLAYERS; ++l) {
        const uint32_t bytes = static_cast<uint32_t>(layer_n(l)) * 128u;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&w_empty[stage], phase ^ 1);
          mbar_expect_tx(&w_full[stage], bytes);
          tma_load_2d(sW + stage * W_STAGE_BYTES, &tm.w[l], &w_full[stage], kb * 64, 0);
          if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // layer:        0 cls      1 logits   2 reg0     3 reg1     4 reg2     5 deltas
    // A operand:    buf0(fc)   buf1       buf0(fc)   buf1       buf0       buf1
    // needs epilogues completed before issue (their output tile and/or the accumulator columns):
    // (layer l waits for all l earlier epilogues: this also keeps acc_full from completing two phases ahead of a waiter)
    const int need[NLAYERS] = {0, 1, 2, 3, 4, 5};
    int stage = 0;
    uint32_t phase = 0;
    int epi_seen = 0;
    mbar_wait(a_full, 0);
    for (int l = 0; l < NLAYERS; ++l) {
      while (epi_seen < need[l]) {
        mbar_wait(epi_done, epi_seen & 1);
        ++epi_seen;
      }
      tc_fence_after();
      const int n = layer_n(l);
      const uint32_t idesc = umma_idesc_f16(TM, n);
      const uint32_t d_tmem = tmem_base + (l == 1 ? COL_LOGIT : (l == 5 ? COL_DELTA : COL_ACC));
      const uint8_t* abuf = sAct + (l & 1) * A_BUF_BYTES;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&w_full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(abuf + kb * A_KB_BYTES));
          const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(sW + stage * W_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&w_empty[stage]);
          if (kb == KB - 1) umma_commit(acc_full);
        }
        __syncwarp();
        if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: two threads per row (column halves) =====================
    const int ew = warp & 3;                  // TMEM sub-partition of this warp
    const int ch = (warp - 4) >> 2;           // column half: 0 -> 0..127, 1 -> 128..255
    const int r = ew * 32 + lane;             // row inside the tile == TMEM lane
    const int t = threadIdx.x - 128;          // 0..255
    const long grow = static_cast<long>(m0) + r;
    const bool valid = grow < p.M;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    // LayerNorm affine parameters of a LN layer -> smem (one float of each per thread)
    auto stage_ln = [&](int idx) {
      sG[t] = __ldg(p.ln_g[idx] + t);
      sBt[t] = __ldg(p.ln_b[idx] + t);
    };
    stage_ln(0);
    int ln_idx = 0;
    for (int l = 0; l < NLAYERS; ++l) {
      mbar_wait(acc_full, l & 1);
      tc_fence_after();
      if (l == 1) {
        // class_logits: 30 of 32 accumulator columns + bias -> fp32 logits
        if (ch == 0) {
          uint32_t v[32];
          tmem_ld32(tlane + COL_LOGIT, v);
          tmem_ld_wait();
          if (valid) {
            float* o = p.logits_out + grow * p.C;
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c < p.C) o[c] = __uint_as_float(v[c]) + __ldg(p.cls_bias + c);
          }
        }
        tc_fence_before();
        mbar_arrive(epi_done);
        continue;
      }
      if (l == 5) {
        // bboxes_delta + apply_deltas (box_head.py:550-590), same arithmetic as head_final_kernel
        if (ch == 0) {
          uint32_t v[32];
          tmem_ld32(tlane + COL_DELTA, v);     // columns 288..319: the first 4 are the deltas
          tmem_ld_wait();
          if (valid) {
            const float4 b = *reinterpret_cast<const float4*>(p.boxes_in + grow * 4);
            const float d0 = __uint_as_float(v[0]) + __ldg(p.delta_bias + 0);
            const float d1 = __uint_as_float(v[1]) + __ldg(p.delta_bias + 1);
            const float d2 = __uint_as_float(v[2]) + __ldg(p.delta_bias + 2);
            const float d3 = __uint_as_float(v[3]) + __ldg(p.delta_bias + 3);
            const float w = __fsub_rn(b.z, b.x), h = __fsub_rn(b.w, b.y);
            const float cx = __fadd_rn(b.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(b.y, __fmul_rn(0.5f, h));
            const float dx = __fdiv_rn(d0, 2.0f), dy = __fdiv_rn(d1, 2.0f);
            const float dw = fminf(d2, kScaleClamp), dh = fminf(d3, kScaleClamp);
            const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
            const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
            float4 o;
            o.x = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
            o.y = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
            o.z = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
            o.w = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
            *reinterpret_cast<float4*>(p.boxes_out + grow * 4) = o;
          }
        }
        continue;
      }
      // ---- LayerNorm(256) + ReLU: this thread owns columns [128 ch, 128 ch + 128) of row r
      named_bar_sync(1, EPI_THREADS);         // gamma / beta of this layer are staged
      const uint32_t tcol = tlane + COL_ACC + ch * 128;
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t v0[32], v1[32];
        tmem_ld32(tcol + c2 * 64, v0);
        tmem_ld32(tcol + c2 * 64 + 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float a = __uint_as_float(v0[j]), b = __uint_as_float(v1[j]);
          sum += a + b;
          sq = fmaf(a, a, sq);
          sq = fmaf(b, b, sq);
        }
      }
      sStat[ch * TM + r] = make_float2(sum, sq);
      named_bar_sync(2, EPI_THREADS);
      const float2 other = sStat[(ch ^ 1) * TM + r];
      const float mean = (sum + other.x) * (1.f / D);
      const float var = fmaxf((sq + other.y) * (1.f / D) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
      uint8_t* dst = sAct + ((l & 1) ^ 1) * A_BUF_BYTES;      // the other activation buffer
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t vv[2][32];
        tmem_ld32(tcol + c2 * 64, vv[0]);
        tmem_ld32(tcol + c2 * 64 + 32, vv[1]);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = ch * 4 + c2 * 2 + h;             // 32-column piece of the 256-wide row
          uint8_t* rowp = dst + (c >> 1) * A_KB_BYTES + r * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; e += 4) {
              const int col = c * 32 + q * 8 + e;
              const float4 g4 = *reinterpret_cast<const float4*>(sG + col);
              const float4 b4 = *reinterpret_cast<const float4*>(sBt + col);
              y[e] = fmaxf((__uint_as_float(vv[h][q * 8 + e]) - mean) * rstd * g4.x + b4.x, 0.f);
              y[e + 1] = fmaxf((__uint_as_float(vv[h][q * 8 + e + 1]) - mean) * rstd * g4.y + b4.y, 0.f);
              y[e + 2] = fmaxf((__uint_as_float(vv[h][q * 8 + e + 2]) - mean) * rstd * g4.z + b4.z, 0.f);
              y[e + 3] = fmaxf((__uint_as_float(vv[h][q * 8 + e + 3]) - mean) * rstd * g4.w + b4.w, 0.f);
            }
            uint4 pk;
            pk.x = pack_half2(y[0], y[1]);
            pk.y = pack_half2(y[2], y[3]);
            pk.z = pack_half2(y[4], y[5]);
            pk.w = pack_half2(y[6], y[7]);
            const int chunk = (c & 1) * 4 + q;          // 16-byte chunk inside the 128-byte (64-column) row
            *reinterpret_cast<uint4*>(rowp + ((chunk ^ (r & 7)) << 4)) = pk;
          }
        }
      }
      fence_proxy_async_smem();               // the next layer's MMA reads this tile through the async proxy
      tc_fence_before();
      ++ln_idx;
      if (ln_idx < 4) {
        named_bar_sync(1, EPI_THREADS);       // everybody is done with this layer's gamma / beta and statistics
        stage_ln(ln_idx);
      }
      mbar_arrive(epi_done);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

// fc [M][256] fp16; weights fp16 K-major: cls_w/reg_w* [256][256], logit_w [32][256] (rows >= C zero),
// delta_w [16][256] (rows >= 4 zero); LayerNorm params fp32 [256]; boxes_in [M][4]; outputs fp32.
int head_tail_launch(const void* fc, const void* cls_w, const float* cls_g, const float* cls_b, const void* logit_w,
                     const float* logit_bias, int C, const void* const* reg_w, const float* const* reg_g,
                     const float* const* reg_b, const void* delta_w, const float* delta_bias, const float* boxes_in,
                     float* logits_out, float* boxes_out, int M, cudaStream_t stream) {
  if (M <= 0 || C <= 0 || C > NL_PAD) return DVID_ERR_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(head_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
      return DVID_ERR_CUDA;
    attr_set = true;
  }
  TailMaps tm;
  {
    const uint64_t dims[2] = {D, static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {D * 2};
    const uint32_t box[2] = {64, TM};
    int r = make_tmap_f16(&tm.a, fc, 2, dims, strides, box, nullptr);
    if (r) return r;
  }
  const void* ws[NLAYERS] = {cls_w, logit_w, reg_w[0], reg_w[1], reg_w[2], delta_w};
  const int ns[NLAYERS] = {D, NL_PAD, D, D, D, ND_PAD};
  for (int l = 0; l < NLAYERS; ++l) {
    const uint64_t dims[2] = {D, static_cast<uint64_t>(ns[l])};
    const uint64_t strides[1] = {D * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(ns[l])};
    int r = make_tmap_f16(&tm.w[l], ws[l], 2, dims, strides, box, nullptr);
    if (r) return r;
  }
  TailParams p;
  p.M = M; p.C = C;
  p.ln_g[0] = cls_g; p.ln_b[0] = cls_b;
  for (int i = 0; i < 3; ++i) { p.ln_g[1 + i] = reg_g[i]; p.ln_b[1 + i] = reg_b[i]; }
  p.cls_bias = logit_bias; p.delta_bias = delta_bias; p.boxes_in = boxes_in;
  p.logits_out = logits_out; p.boxes_out = boxes_out;
  for (int l = 0; l < NLAYERS; ++l) p.w_ptr[l] = static_cast<const uint8_t*>(ws[l]);
  launch_pdl(head_tail_kernel, dim3((M + TM - 1) / TM), dim3(384), SMEM_BYTES, stream, tm, p);
  return check_launch();
}

}  // namespace dvid
