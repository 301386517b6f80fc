// Post-processing and memory-management kernels of the DiffusionVID hot path, one CTA per frame / per problem:
//   * top-k over the N x C sigmoid scores          (mega_core/modeling/detector/diffusion_det.py:772-784)
//   * class-aware NMS (torchvision batched_nms coordinate-offset trick, diffusion_det.py:617,793) with the greedy sweep
//     done on the device (the reference's own mega_core/csrc/cuda/nms.cu:100-123 copies the mask to the host), plus
//     the legacy "+1 pixel" IoU variant of mega_core/csrc/{cpu/nms_cpu.cpp,cuda/nms.cu} behind the same kernel
//   * pairwise L2 distances + farthest-point sampling (diffusion_det.py:880-895, mega_core/csrc/cuda/fps.cu:25-142)
// Index work is bit-exact by construction: all float comparisons use the reference's operation order with
// contraction disabled (__f*_rn), ties are broken by ascending index.
#include "dvid_internal.h"
#include "warp_mma.cuh"

namespace dvid {

namespace {

// In-place bitonic sort (descending) of n (power of two) 64-bit keys in shared memory by the whole CTA.
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) {
            keys[i] = b;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ float sigmoidf_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// ------------------------------------------------------------------------------------------------ top-k of N*C scores
// Exclusive prefix sum of one int per thread over the CTA (1024 threads); returns the prefix, *total = CTA sum.
__device__ __forceinline__ int block_excl_scan(int v, int* swarp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) swarp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = swarp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    swarp[lane] = w;      // inclusive over warps
  }
  __syncthreads();
  const int base = warp ? swarp[warp - 1] : 0;
  *total = swarp[31];
  __syncthreads();        // swarp may be reused by the caller
  return base + inc - v;
}

// Per frame: radix-select the k-th largest score (4 x 8-bit passes over the fp32 bit patterns - sigmoid outputs are
// positive, so they order like unsigned ints), compact the k winners in flat-index order (ties at the threshold are
// resolved towards the lowest flat index) and bitonic-sort just those <= 1024 keys.
// key = score bits << 32 | ~flat_index -> descending sort gives score-descending, flat-index-ascending order (the
// canonical order of SURVEY.md 8c contract 3).  The first version sorted all 16384 padded keys (255 us per call).
__global__ void __launch_bounds__(1024)
topk_scores_kernel(const float* __restrict__ logits, const float* __restrict__ boxes, int N, int C, int k,
                   float* __restrict__ out_boxes, float* __restrict__ out_scores, int* __restrict__ out_labels,
                   int cap, int slot0) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char tk_smem[];
  unsigned long long* skeys = reinterpret_cast<unsigned long long*>(tk_smem);        // 1024 selected keys
  unsigned* su = reinterpret_cast<unsigned*>(tk_smem + 1024 * 8);                      // N*C score bit patterns
  __shared__ int hist[256];
  __shared__ int swarp[32];
  __shared__ unsigned s_prefix;
  __shared__ int s_remaining;
  const int f = blockIdx.x;
  const int tid = threadIdx.x;
  const int total = N * C;
  const float* lg = logits + static_cast<long>(f) * total;
  for (int i = tid; i < total; i += blockDim.x) su[i] = __float_as_uint(sigmoidf_ref(lg[i]));
  skeys[tid] = 0ull;
  if (tid == 0) { s_prefix = 0u; s_remaining = k; }
  __syncthreads();
  // ---- radix select: after pass p the top (4-p) bytes of the k-th largest value are known
  for (int pass = 3; pass >= 0; --pass) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned hi_mask = pass == 3 ? 0u : (0xFFFFFFFFu << (8 * (pass + 1)));
    for (int i = tid; i < total; i += blockDim.x) {
      const unsigned u = su[i];
      if ((u & hi_mask) == prefix) atomicAdd(&hist[(u >> (8 * pass)) & 255u], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int rem = s_remaining, b = 255;
      for (; b > 0; --b) {
        if (hist[b] >= rem) break;
        rem -= hist[b];
      }
      s_prefix = prefix | (static_cast<unsigned>(b) << (8 * pass));
      s_remaining = rem;     // how many elements of the chosen bin (finally: equal to the threshold) are still needed
    }
    __syncthreads();
  }
  const unsigned thr = s_prefix;
  const int need_eq = s_remaining;
  // ---- ordered compaction: thread t owns the contiguous flat indices [t*per, (t+1)*per)
  const int per = (total + 1023) / 1024;
  const int i0 = tid * per, i1 = min(total, i0 + per);
  int n_gt = 0, n_eq = 0;
  for (int i = i0; i < i1; ++i) {
    const unsigned u = su[i];
    n_gt += (u > thr);
    n_eq += (u == thr);
  }
  int tot_gt, tot_eq;
  int p_gt = block_excl_scan(n_gt, swarp, &tot_gt);
  int p_eq = block_excl_scan(n_eq, swarp, &tot_eq);
  for (int i = i0; i < i1; ++i) {
    const unsigned u = su[i];
    int slot = -1;
    if (u > thr) slot = p_gt++;
    else if (u == thr) {
      if (p_eq < need_eq) slot = tot_gt + p_eq;
      ++p_eq;
    }
    if (slot >= 0) skeys[slot] = (static_cast<unsigned long long>(u) << 32) | (0xFFFFFFFFu - static_cast<unsigned>(i));
  }
  __syncthreads();
  bitonic_sort_desc(skeys, 1024);
  if (tid < k) {
    const unsigned long long key = skeys[tid];
    const unsigned idx = 0xFFFFFFFFu - static_cast<unsigned>(key & 0xFFFFFFFFull);
    const int box = idx / C, cls = idx % C;
    const long o = static_cast<long>(f) * cap + slot0 + tid;
    out_scores[o] = __uint_as_float(static_cast<unsigned>(key >> 32));
    out_labels[o] = cls + 1;
    *reinterpret_cast<float4*>(out_boxes + o * 4) =
        *reinterpret_cast<const float4*>(boxes + (static_cast<long>(f) * N + box) * 4);
  }
}

// per-frame top-k of the max logit -> 0/1 mask in box order (box_head.py:304-311); k1 >= k2, mask2 = first k2 of k1
__global__ void __launch_bounds__(1024)
topk_mask_kernel(const float* __restrict__ logits, int N, int C, int k1, int k2, unsigned char* __restrict__ mask1,
                 unsigned char* __restrict__ mask2) {
  pdl_prologue();
  __shared__ unsigned long long skeys[1024];
  const int f = blockIdx.x;
  const int i = threadIdx.x;
  unsigned long long key = 0ull;
  if (i < N) {
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, logits[(static_cast<long>(f) * N + i) * C + c]);
    unsigned u = __float_as_uint(mx);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // order-preserving map of signed floats to unsigned
    key = (static_cast<unsigned long long>(u) << 32) | (0xFFFFFFFFu - static_cast<unsigned>(i));
    mask1[static_cast<long>(f) * N + i] = 0;
    mask2[static_cast<long>(f) * N + i] = 0;
  }
  skeys[i] = key;
  __syncthreads();
  bitonic_sort_desc(skeys, 1024);
  if (i < k1) {
    const unsigned idx = 0xFFFFFFFFu - static_cast<unsigned>(skeys[i] & 0xFFFFFFFFull);
    mask1[static_cast<long>(f) * N + idx] = 1;
    if (i < k2) mask2[static_cast<long>(f) * N + idx] = 1;
  }
}

// gather rows of src [frames*N][256] fp32 whose mask is set, in (frame, box) order, `k` rows per frame
__global__ void __launch_bounds__(1024)
gather_masked_rows_kernel(const float* __restrict__ src, const unsigned char* __restrict__ mask, int N, int k,
                          float* __restrict__ dst) {
  pdl_prologue();
  __shared__ int spos[1024];
  __shared__ int wcnt[32];
  const int f = blockIdx.x, i = threadIdx.x, lane = i & 31, warp = i >> 5;
  const bool m = (i < N) && mask[static_cast<long>(f) * N + i];
  const unsigned ball = __ballot_sync(0xffffffffu, m);
  if (lane == 0) wcnt[warp] = __popc(ball);
  __syncthreads();
  int off = 0;
  for (int w = 0; w < warp; ++w) off += wcnt[w];
  spos[i] = m ? off + __popc(ball & ((1u << lane) - 1u)) : -1;
  __syncthreads();
  for (int r = warp; r < N; r += 32) {
    const int p = spos[r];
    if (p >= 0 && p < k) {
      const float* s = src + (static_cast<long>(f) * N + r) * 256;
      float* d = dst + (static_cast<long>(f) * k + p) * 256;
      *reinterpret_cast<float4*>(d + lane * 8) = *reinterpret_cast<const float4*>(s + lane * 8);
      *reinterpret_cast<float4*>(d + lane * 8 + 4) = *reinterpret_cast<const float4*>(s + lane * 8 + 4);
    }
  }
}

// ------------------------------------------------------------------------------------------------ NMS
struct NmsArgs {
  const float* boxes;    // [frames][cap][4]
  const float* scores;   // [frames][cap]
  const int* labels;     // [frames][cap] or nullptr (class-agnostic)
  const int* counts;     // [frames] or nullptr (then n)
  int n, cap;
  float thr;
  int plus_one;          // legacy +1 pixel IoU (mega_core/csrc/cuda/nms.cu:13-21)
  int ge;                // suppress when IoU >= thr (mega_core/csrc/cpu/nms_cpu.cpp:58) instead of >
  int ascending_out;     // output kept indices in ascending index order (legacy _C.nms) instead of score order
  float clip_w, clip_h;  // BoxList.clip_to_image(TO_REMOVE=1) applied to out_boxes when > 0
  long long* keep_idx;   // [frames][cap] int64 or nullptr
  float* out_boxes; float* out_scores; int* out_labels;   // compacted outputs or nullptr
  int* out_count;        // [frames]
  // multi-kernel path (workspace given): phase 1 = sort, (mask kernel), phase 2 = sweep + outputs; phase 0 = all in one
  int phase;
  unsigned long long* ws_keys;   // [frames][1024]
  float4* ws_boxes;              // [frames][1024] sorted, class-offset boxes
  unsigned long long* ws_mask;   // [frames][1024][16]
};

constexpr int NMS_MAX = 1024;
constexpr int NMS_WORDS = NMS_MAX / 64;

__global__ void __launch_bounds__(1024) nms_kernel(const NmsArgs a) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char nms_smem[];
  unsigned long long* skeys = reinterpret_cast<unsigned long long*>(nms_smem);             // 1024 keys
  float4* sbox = reinterpret_cast<float4*>(nms_smem + NMS_MAX * 8);                        // 1024 sorted boxes
  unsigned long long* smask = reinterpret_cast<unsigned long long*>(nms_smem + NMS_MAX * 24);  // [1024][16]
  __shared__ float sred[32];
  __shared__ unsigned long long skept[NMS_WORDS];   // over sorted positions
  __shared__ unsigned long long skept2[NMS_WORDS];  // over original indices
  __shared__ int sprefix[NMS_WORDS + 1];

  const int f = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.counts ? min(a.counts[f], a.cap) : a.n;
  const float* boxes = a.boxes + static_cast<long>(f) * a.cap * 4;
  const float* scores = a.scores + static_cast<long>(f) * a.cap;
  const int* labels = a.labels ? a.labels + static_cast<long>(f) * a.cap : nullptr;

  const int words = (n + 63) >> 6;
  const float one = a.plus_one ? 1.0f : 0.0f;
  if (a.phase == 2) {
    // sorted keys and the suppression matrix were produced by phase 1 + nms_mask_kernel
    skeys[tid] = a.ws_keys[static_cast<long>(f) * NMS_MAX + tid];
    if (tid < NMS_WORDS) { skept[tid] = 0ull; skept2[tid] = 0ull; }
    const uint4* src = reinterpret_cast<const uint4*>(a.ws_mask + static_cast<long>(f) * NMS_MAX * NMS_WORDS);
    uint4* dst = reinterpret_cast<uint4*>(smask);
    for (int i = tid; i < n * NMS_WORDS / 2; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
  } else {
  // max coordinate (torchvision batched_nms: offsets = idxs * (boxes.max() + 1))
  float mx = -INFINITY;
  float4 mybox = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid < n) {
    mybox = *reinterpret_cast<const float4*>(boxes + tid * 4);
    mx = fmaxf(fmaxf(mybox.x, mybox.y), fmaxf(mybox.z, mybox.w));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) sred[warp] = mx;
  if (tid < NMS_WORDS) { skept[tid] = 0ull; skept2[tid] = 0ull; }
  __syncthreads();
  mx = sred[0];
  for (int w = 1; w < 32; ++w) mx = fmaxf(mx, sred[w]);

  unsigned long long key = 0ull;
  if (tid < n) {
    unsigned u = __float_as_uint(scores[tid]);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    key = (static_cast<unsigned long long>(u) << 32) | (0xFFFFFFFFu - static_cast<unsigned>(tid));
  }
  skeys[tid] = key;
  __syncthreads();
  bitonic_sort_desc(skeys, NMS_MAX);

  if (tid < n) {
    const int src = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(skeys[tid] & 0xFFFFFFFFull));
    float4 b = *reinterpret_cast<const float4*>(boxes + src * 4);
    if (labels) {
      const float off = __fmul_rn(static_cast<float>(labels[src]), __fadd_rn(mx, 1.0f));
      b.x = __fadd_rn(b.x, off); b.y = __fadd_rn(b.y, off); b.z = __fadd_rn(b.z, off); b.w = __fadd_rn(b.w, off);
    }
    sbox[tid] = b;
  }
  __syncthreads();

  if (a.phase == 1) {
    a.ws_keys[static_cast<long>(f) * NMS_MAX + tid] = skeys[tid];
    a.ws_boxes[static_cast<long>(f) * NMS_MAX + tid] = (tid < n) ? sbox[tid] : make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }

  // suppression bit matrix over sorted positions: bit j of smask[i][w] set iff j > i and IoU(i, j) beats thr.
  // One warp per row i, lanes over consecutive j (conflict-free sbox reads, bi is a broadcast), 32 bits per ballot.
  unsigned* smask32 = reinterpret_cast<unsigned*>(smask);
  for (int i = warp; i < n; i += 32) {
    const float4 bi = sbox[i];
    const float sa = __fmul_rn(__fadd_rn(__fsub_rn(bi.z, bi.x), one), __fadd_rn(__fsub_rn(bi.w, bi.y), one));
    for (int h = 0; h < 2 * words; ++h) {          // 32-bit half words
      const int j = h * 32 + lane;
      bool sup = false;
      if (h * 32 + 31 > i && j > i && j < n) {
        const float4 bj = sbox[j];
        const float left = fmaxf(bi.x, bj.x), right = fminf(bi.z, bj.z);
        const float top = fmaxf(bi.y, bj.y), bottom = fminf(bi.w, bj.w);
        const float width = fmaxf(__fadd_rn(__fsub_rn(right, left), one), 0.f);
        const float height = fmaxf(__fadd_rn(__fsub_rn(bottom, top), one), 0.f);
        const float inter = __fmul_rn(width, height);
        // disjoint boxes (almost all pairs: other classes sit at other coordinate offsets): IoU = 0 (or 0/0 = NaN),
        // neither beats a positive threshold - skip the division.  Exact: same outcome as evaluating it.
        if (inter > 0.f || !(a.thr > 0.f)) {
          const float sb = __fmul_rn(__fadd_rn(__fsub_rn(bj.z, bj.x), one), __fadd_rn(__fsub_rn(bj.w, bj.y), one));
          const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter));
          sup = a.ge ? (iou >= a.thr) : (iou > a.thr);
        }
      }
      const unsigned bits = __ballot_sync(0xffffffffu, sup);
      if (lane == 0) smask32[(i * NMS_WORDS) * 2 + h] = bits;      // little-endian halves of the 64-bit word
    }
  }
  __syncthreads();
  }   // phase != 2

  // greedy sweep by warp 0, 64 sorted boxes at a time: the chunk's own (diagonal) words resolve the chunk serially
  // (one dependent smem read per box), then the surviving rows are OR-ed into the removed set, lane w owning word w.
  if (warp == 0) {
    unsigned long long removed = 0ull, kept = 0ull;
    for (int c = 0; c < words; ++c) {
      unsigned long long cur = __shfl_sync(0xffffffffu, removed, c);
      unsigned long long kc = 0ull;
      const int cn = min(64, n - c * 64);
      for (int bb = 0; bb < cn; ++bb) {
        if (!((cur >> bb) & 1ull)) {
          kc |= (1ull << bb);
          cur |= smask[(c * 64 + bb) * NMS_WORDS + c];
        }
      }
      if (lane == c) kept = kc;
      if (lane > c && lane < words) {
        unsigned long long acc = removed;
        unsigned long long m = kc;
        while (m) {
          const int bb = __ffsll(static_cast<long long>(m)) - 1;
          m &= m - 1;
          acc |= smask[(c * 64 + bb) * NMS_WORDS + lane];
        }
        removed = acc;
      }
    }
    if (lane < NMS_WORDS) skept[lane] = (lane < words) ? kept : 0ull;
  }
  __syncthreads();

  const bool is_kept = (tid < n) && ((skept[tid >> 6] >> (tid & 63)) & 1ull);
  const int orig = (tid < n) ? static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(skeys[tid] & 0xFFFFFFFFull)) : 0;
  if (a.ascending_out && is_kept) atomicOr(&skept2[orig >> 6], 1ull << (orig & 63));
  __syncthreads();
  const unsigned long long* km = a.ascending_out ? skept2 : skept;
  if (tid == 0) {
    int s = 0;
    for (int w = 0; w < NMS_WORDS; ++w) { sprefix[w] = s; s += __popcll(km[w]); }
    sprefix[NMS_WORDS] = s;
    a.out_count[f] = s;
  }
  __syncthreads();
  if (is_kept) {
    const int pos_bit = a.ascending_out ? orig : tid;
    const int rank = sprefix[pos_bit >> 6] + __popcll(km[pos_bit >> 6] & ((1ull << (pos_bit & 63)) - 1ull));
    const long o = static_cast<long>(f) * a.cap + rank;
    if (a.keep_idx) a.keep_idx[o] = orig;
    if (a.out_boxes) {
      float4 b = *reinterpret_cast<const float4*>(boxes + orig * 4);
      if (a.clip_w > 0.f) {
        b.x = fminf(fmaxf(b.x, 0.f), a.clip_w - 1.0f); b.y = fminf(fmaxf(b.y, 0.f), a.clip_h - 1.0f);
        b.z = fminf(fmaxf(b.z, 0.f), a.clip_w - 1.0f); b.w = fminf(fmaxf(b.w, 0.f), a.clip_h - 1.0f);
      }
      *reinterpret_cast<float4*>(a.out_boxes + o * 4) = b;
    }
    if (a.out_scores) a.out_scores[o] = scores[orig];
    if (a.out_labels && labels) a.out_labels[o] = labels[orig];
  }
}

// Suppression matrix for the multi-kernel NMS path: one warp per sorted row i, grid (ceil(1024/8), frames), so the
// O(n^2) IoU work of a frame spreads over ~113 CTAs instead of one (the single-CTA version was issue-bound at ~140 us).
__global__ void __launch_bounds__(256)
nms_mask_kernel(const float4* __restrict__ ws_boxes, const int* __restrict__ counts, int n_fixed, int cap, float thr,
                int plus_one, int ge, unsigned long long* __restrict__ ws_mask) {
  pdl_prologue();
  const int f = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int n = counts ? min(counts[f], cap) : n_fixed;
  if (i >= n) return;
  const int words = (n + 63) >> 6;
  const float one = plus_one ? 1.0f : 0.0f;
  const float4* sbox = ws_boxes + static_cast<long>(f) * NMS_MAX;
  unsigned* m32 = reinterpret_cast<unsigned*>(ws_mask + (static_cast<long>(f) * NMS_MAX + i) * NMS_WORDS);
  const float4 bi = sbox[i];
  const float sa = __fmul_rn(__fadd_rn(__fsub_rn(bi.z, bi.x), one), __fadd_rn(__fsub_rn(bi.w, bi.y), one));
  for (int h = 0; h < 2 * words; ++h) {
    const int j = h * 32 + lane;
    bool sup = false;
    if (h * 32 + 31 > i && j > i && j < n) {
      const float4 bj = sbox[j];
      const float left = fmaxf(bi.x, bj.x), right = fminf(bi.z, bj.z);
      const float top = fmaxf(bi.y, bj.y), bottom = fminf(bi.w, bj.w);
      const float width = fmaxf(__fadd_rn(__fsub_rn(right, left), one), 0.f);
      const float height = fmaxf(__fadd_rn(__fsub_rn(bottom, top), one), 0.f);
      const float inter = __fmul_rn(width, height);
      if (inter > 0.f || !(thr > 0.f)) {       // see nms_kernel: skipping the division for disjoint boxes is exact
        const float sb = __fmul_rn(__fadd_rn(__fsub_rn(bj.z, bj.x), one), __fadd_rn(__fsub_rn(bj.w, bj.y), one));
        const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter));
        sup = ge ? (iou >= thr) : (iou > thr);
      }
    }
    const unsigned bits = __ballot_sync(0xffffffffu, sup);
    if (lane == 0) m32[h] = bits;
  }
}

// ------------------------------------------------------------------------------------------------ cdist + FPS
// out[i][j] = sqrt(sum_k (x[i][k] - x[j][k])^2), fp32, direct differences (torch.cdist p=2 without the matmul trick).
__global__ void __launch_bounds__(256) cdist_kernel(const float* __restrict__ x, float* __restrict__ out, int n, int d) {
  pdl_prologue();
  __shared__ float sa[32][33];
  __shared__ float sb[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < d; k0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = ty + r * 8;
      sa[row][tx] = (i0 + row < n && k0 + tx < d) ? x[static_cast<long>(i0 + row) * d + k0 + tx] : 0.f;
      sb[row][tx] = (j0 + row < n && k0 + tx < d) ? x[static_cast<long>(j0 + row) * d + k0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float bv = sb[tx][k];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float df = sa[ty + r * 8][k] - bv;
        acc[r] = fmaf(df, df, acc[r]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + r * 8, j = j0 + tx;
    if (i < n && j < n) out[static_cast<long>(i) * n + j] = sqrtf(acc[r]);
  }
}

// Farthest point sampling over a precomputed distance matrix: same contract as mega_core._C.furthest_point_sampling
// (csrc/fps.h:15-36): dist (b,n,n), temp (b,n) running minimum (caller fills 1e10), idx (b,m) int32, idx[0] = 0.
// Tie rule of the reference kernel: thread t = k mod bs scans its k's with strict '>' (lowest k wins inside a thread);
// the shared-memory tree then merges slot t with t+s for s = bs/2 .. 1, the lower slot winning ties.  Two tied slots
// first meet at s = lowest differing bit of their indices and the one with that bit clear wins, i.e. the winner among
// equal maxima has the smallest (bit_reverse(k mod bs), k); bs = the reference's block size for this n.
constexpr int FPS_PER_THREAD = 8;   // n <= 8192
// THREADS: the greedy loop is a chain of m-1 rounds of (row load, block arg-max, broadcast), pure latency.  The memory
// sizes of the path (1800 -> 900 and 600 -> 150 candidates per video) fit 256 threads x 8 points: 8 independent loads
// per thread, a second-level reduction over 8 instead of 32 warps and 256-thread barriers (1.2 -> 0.8 us per round on
// B200).  The pick does not depend on the block size - ties are resolved by the reference block's priority (above).
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
fps_kernel(int n, int m, int log2_bs, const float* __restrict__ dist, float* __restrict__ temp,
           int* __restrict__ idx) {
  constexpr int NWARPS = THREADS / 32;
  pdl_prologue();
  __shared__ float sval[32];
  __shared__ unsigned sprio[32];
  __shared__ int sold;
  const int batch = blockIdx.x;
  dist += static_cast<long>(batch) * n * n;
  temp += static_cast<long>(batch) * n;
  idx += static_cast<long>(batch) * m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float tv[FPS_PER_THREAD];
#pragma unroll
  for (int r = 0; r < FPS_PER_THREAD; ++r) {
    const int k = tid + r * THREADS;
    tv[r] = k < n ? temp[k] : 0.f;
  }
  int old = 0;
  if (tid == 0 && m > 0) idx[0] = 0;
  const unsigned slot_bits = 16;   // priority = bitrev(k % bs) << 16 | k  (n <= 8192 < 2^16); smaller wins
  const unsigned bs_mask = (1u << log2_bs) - 1u;
  for (int j = 1; j < m; ++j) {
    float best = -1.f;
    unsigned bprio = 0xFFFFFFFFu;
    const float* row = dist + static_cast<long>(old) * n;
    float rv[FPS_PER_THREAD];
#pragma unroll
    for (int r = 0; r < FPS_PER_THREAD; ++r) {      // all loads of the round in flight before the first compare
      const int k = tid + r * THREADS;
      rv[r] = k < n ? __ldg(row + k) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < FPS_PER_THREAD; ++r) {
      const int k = tid + r * THREADS;
      if (k < n) {
        const float d2 = fminf(rv[r], tv[r]);
        tv[r] = d2;
        const unsigned slot = static_cast<unsigned>(k) & bs_mask;
        const unsigned rev = log2_bs ? (__brev(slot) >> (32 - log2_bs)) : 0u;
        const unsigned pr = (rev << slot_bits) | static_cast<unsigned>(k);
        if (d2 > best || (d2 == best && pr < bprio)) { best = d2; bprio = pr; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const unsigned op = __shfl_xor_sync(0xffffffffu, bprio, o);
      if (ov > best || (ov == best && op < bprio)) { best = ov; bprio = op; }
    }
    if (lane == 0) { sval[warp] = best; sprio[warp] = bprio; }
    __syncthreads();
    if (warp == 0) {
      best = lane < NWARPS ? sval[lane] : -2.f;
      bprio = lane < NWARPS ? sprio[lane] : 0xFFFFFFFFu;
#pragma unroll
      for (int o = NWARPS / 2; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const unsigned op = __shfl_xor_sync(0xffffffffu, bprio, o);
        if (ov > best || (ov == best && op < bprio)) { best = ov; bprio = op; }
      }
      if (lane == 0) {
        // the reference starts every thread at (best=-1, besti=0): if nothing beats -1 the pick is index 0
        const int pick = (best > -1.f) ? static_cast<int>(bprio & 0xFFFFu) : 0;
        sold = pick;
        idx[j] = pick;
      }
    }
    __syncthreads();
    old = sold;
  }
#pragma unroll
  for (int r = 0; r < FPS_PER_THREAD; ++r) {
    const int k = tid + r * THREADS;
    if (k < n) temp[k] = tv[r];
  }
}

}  // namespace

int topk_scores_launch(const float* logits, const float* boxes, int frames, int N, int C, int k, float* out_boxes,
                       float* out_scores, int* out_labels, int cap, int slot0, cudaStream_t stream) {
  if (frames <= 0 || N <= 0 || C <= 0 || k <= 0 || k > N * C || k > 1024 || slot0 + k > cap) return DVID_ERR_SHAPE;
  const size_t smem = 1024 * 8 + static_cast<size_t>(N) * C * 4;
  if (smem > 200 * 1024) return DVID_ERR_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(topk_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
        cudaSuccess)
      return DVID_ERR_CUDA;
    attr_set = true;
  }
  launch_pdl(topk_scores_kernel, dim3(frames), dim3(1024), smem, stream, logits, boxes, N, C, k, out_boxes, out_scores, out_labels, cap,
                                                     slot0);
  return check_launch();
}

int topk_mask_launch(const float* logits, int frames, int N, int C, int k1, int k2, unsigned char* mask1,
                     unsigned char* mask2, cudaStream_t stream) {
  if (frames <= 0 || N <= 0 || N > 1024 || k1 > N || k2 > k1 || k2 < 0) return DVID_ERR_SHAPE;
  launch_pdl(topk_mask_kernel, dim3(frames), dim3(1024), 0, stream, logits, N, C, k1, k2, mask1, mask2);
  return check_launch();
}

int gather_masked_rows_launch(const float* src, const unsigned char* mask, int frames, int N, int k, float* dst,
                              cudaStream_t stream) {
  if (frames <= 0 || N <= 0 || N > 1024 || k <= 0) return DVID_ERR_SHAPE;
  launch_pdl(gather_masked_rows_kernel, dim3(frames), dim3(1024), 0, stream, src, mask, N, k, dst);
  return check_launch();
}

int nms_launch(const float* boxes, const float* scores, const int* labels, const int* counts, int n, int cap,
               int frames, float thr, int plus_one, int ge, int ascending_out, float clip_w, float clip_h,
               long long* keep_idx, float* out_boxes, float* out_scores, int* out_labels, int* out_count,
               void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (frames <= 0 || cap <= 0 || n < 0 || n > NMS_MAX || n > cap || out_count == nullptr) return DVID_ERR_SHAPE;
  if (counts != nullptr && cap > NMS_MAX) return DVID_ERR_SHAPE;
  const int smem = NMS_MAX * 24 + NMS_MAX * NMS_WORDS * 8;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return DVID_ERR_CUDA;
    attr_set = true;
  }
  NmsArgs a;
  a.boxes = boxes; a.scores = scores; a.labels = labels; a.counts = counts; a.n = n; a.cap = cap; a.thr = thr;
  a.plus_one = plus_one; a.ge = ge; a.ascending_out = ascending_out; a.clip_w = clip_w; a.clip_h = clip_h;
  a.keep_idx = keep_idx; a.out_boxes = out_boxes; a.out_scores = out_scores; a.out_labels = out_labels;
  a.out_count = out_count;
  a.phase = 0; a.ws_keys = nullptr; a.ws_boxes = nullptr; a.ws_mask = nullptr;
  const size_t per_frame = NMS_MAX * 8 + NMS_MAX * 16 + static_cast<size_t>(NMS_MAX) * NMS_WORDS * 8;
  if (workspace != nullptr && workspace_bytes >= per_frame * frames) {
    // sort -> suppression matrix over many CTAs -> sweep + outputs
    unsigned char* w = static_cast<unsigned char*>(workspace);
    a.ws_keys = reinterpret_cast<unsigned long long*>(w);
    a.ws_boxes = reinterpret_cast<float4*>(w + static_cast<size_t>(frames) * NMS_MAX * 8);
    a.ws_mask = reinterpret_cast<unsigned long long*>(w + static_cast<size_t>(frames) * NMS_MAX * 24);
    a.phase = 1;
    launch_pdl(nms_kernel, dim3(frames), dim3(1024), smem, stream, a);
    dim3 grid((NMS_MAX + 7) / 8, frames);
    launch_pdl(nms_mask_kernel, dim3(grid), dim3(256), 0, stream, a.ws_boxes, counts, n, cap, thr, plus_one, ge, a.ws_mask);
    a.phase = 2;
    launch_pdl(nms_kernel, dim3(frames), dim3(1024), smem, stream, a);
    return check_launch();
  }
  launch_pdl(nms_kernel, dim3(frames), dim3(1024), smem, stream, a);
  return check_launch();
}

int cdist_launch(const float* x, float* out, int n, int d, cudaStream_t stream) {
  if (n <= 0 || d <= 0) return DVID_ERR_SHAPE;
  dim3 grid((n + 31) / 32, (n + 31) / 32);
  launch_pdl(cdist_kernel, dim3(grid), dim3(256), 0, stream, x, out, n, d);
  return check_launch();
}

int fps_launch(int b, int n, int m, const float* dist, float* temp, int* idx, cudaStream_t stream) {
  if (b <= 0 || n <= 0 || m < 0 || m > n || n > 1024 * FPS_PER_THREAD) return DVID_ERR_SHAPE;
  if (m == 0) return DVID_OK;
  int p = 0;
  while ((2 << p) <= n) ++p;           // floor(log2 n): opt_n_threads of the reference (fps.cu:11-15)
  if (p > 10) p = 10;
  if (n <= 256 * FPS_PER_THREAD)
    launch_pdl(fps_kernel<256>, dim3(b), dim3(256), 0, stream, n, m, p, dist, temp, idx);
  else
    launch_pdl(fps_kernel<1024>, dim3(b), dim3(1024), 0, stream, n, m, p, dist, temp, idx);
  return check_launch();
}

}  // namespace dvid
