// K3a: stable LSD radix sort of (id, position) pairs -- the "sort" half of the sort/segment-reduce that
// replaces TF optimizer.py::_deduplicate_indexed_slices (tf.unique + tf.unsorted_segment_sum; SURVEY
// A.3, reached from ops.py:144/148 through Optimizer.minimize).
//
// One launch sorts TWO independent key arrays (the batch's user ids and item ids): the
// first half of the grid owns problem A, the second half problem B, and both advance through the
// digit passes (<= 9 bits each: 2 passes for the 18-bit ids of ML-25M) in lock step with grid-wide barriers.  Per pass: (a) per-CTA digit histogram of
// the CTA's contiguous chunk, (b) grid barrier, (c) every CTA derives its own scatter bases from all
// histograms (digit-exclusive scan + counts of lower CTAs), (d) stable scatter tile by tile: inside a
// warp equal digits are ranked with __match_any_sync, across warps by a per-digit walk over the 16
// warps' counters.  Because positions start as 0..n-1 and every pass is stable, equal ids end up in
// ascending position = batch order, which is what makes the segment sums reproduce
// unsorted_segment_sum's in-order adds.
#include "common.cuh"

namespace tfr {

// Grid-wide barrier of a grid whose CTAs all become resident (the grid is sized against the device's resident
// capacity at launch, sort_resident_capacity: <= 128 CTAs of 256 threads on 148 SMs; and no kernel that can run beside
// this one waits on it).  A plain launch + this barrier instead of a cooperative
// launch: the driver gang-schedules cooperative grids only onto an otherwise idle GPU, which kept the next
// batch's sort from running under the current step's table pass (measured with tools/timeline.py).
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
}

constexpr int SORT_THREADS = 256;  // 8 warps, ~8K registers per CTA: fits beside the table pass's 3 x 256 x 64
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int RADIX_BITS_MAX = 9;           // digit width is chosen per call: ceil(key bits / passes) <= 9
constexpr int RADIX = 1 << RADIX_BITS_MAX;  // 512 bins -> 18-bit ids (ML-25M) in 2 passes, 27-bit (1e8) in 3
constexpr int SORT_MAX_BPP = 64;  // CTAs per problem; 2*64 <= 148 SMs keeps the grid co-resident

struct SortProblem {
  const int32_t* in_ids;
  int32_t* out_ids;
  int32_t* out_pos;
  int32_t* tmp_ids;
  int32_t* tmp_pos;
  int64_t n;
};

// SPLIT = false: the whole sort in ONE launch, the phases separated by grid barriers (every CTA must be resident).
// SPLIT = true: ONE phase per launch -- (only_pass, only_part 0) the histograms of a pass, (only_pass, 1) its scan +
// scatter; the kernel boundaries are the barriers, and no CTA ever waits for another: the form that runs beside a kernel
// that fills the machine without the all-resident requirement (the host-fed step's graph, capi.cu).  Same code, same result.
// (8 CTAs per SM = 32 registers per thread: a CTA's 8K registers are exactly what two CTAs of the table pass leave)
template <bool SPLIT>
__global__ void __launch_bounds__(SORT_THREADS, 8) dedup_sort_kernel(SortProblem pa, SortProblem pb, int bpp, int n_passes,
                                                                  int digit_bits, uint32_t* __restrict__ hist_g,
                                                                  const tfr_opt_scalars* __restrict__ opt,
                                                                  unsigned int* __restrict__ barrier, int only_pass,
                                                                  int only_part) {
  TlScope tl_scope(opt, TFR_TL_SORT);
  unsigned int epoch = 0;
  __shared__ uint32_t s_wc[SORT_WARPS][RADIX];
  __shared__ uint32_t s_base[RADIX];
  __shared__ uint32_t s_scan[SORT_WARPS];
  static_assert(RADIX % SORT_THREADS == 0, "whole digits per thread");

  const int prob = blockIdx.x / bpp;
  const int blk = blockIdx.x % bpp;
  const SortProblem p = prob ? pb : pa;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t dmask = (1u << digit_bits) - 1u;
  const uint32_t invalid_digit = RADIX;  // never matches a real digit in __match_any_sync

  int64_t chunk = (p.n + bpp - 1) / bpp;
  chunk = (chunk + SORT_THREADS - 1) / SORT_THREADS * SORT_THREADS;
  const int64_t begin = min((int64_t)blk * chunk, p.n);
  const int64_t end = min(begin + chunk, p.n);
  uint32_t* my_hist = hist_g + ((size_t)prob * bpp + blk) * RADIX;

  for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&s_wc[0][0])[i] = 0;

  for (int pass = 0; pass < n_passes; ++pass) {
    if (SPLIT && pass != only_pass) continue;
    const int shift = pass * digit_bits;
    // ping-pong so that the LAST pass lands in out_*
    const bool to_out = ((n_passes - 1 - pass) & 1) == 0;
    const int32_t* src_ids = pass == 0 ? p.in_ids : (to_out ? p.tmp_ids : p.out_ids);
    const int32_t* src_pos = pass == 0 ? nullptr : (to_out ? p.tmp_pos : p.out_pos);
    int32_t* dst_ids = to_out ? p.out_ids : p.tmp_ids;
    int32_t* dst_pos = to_out ? p.out_pos : p.tmp_pos;

    // (a) histogram of my chunk
    if (!SPLIT || only_part == 0) {
      for (int dg = tid; dg < RADIX; dg += SORT_THREADS) s_base[dg] = 0;
      __syncthreads();
      for (int64_t i0 = begin; i0 < end; i0 += SORT_THREADS) {
        const int64_t i = i0 + tid;
        const bool valid = i < end;
        const uint32_t digit = valid ? (((uint32_t)src_ids[i] >> shift) & dmask) : invalid_digit;
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        if (valid && (peers & lt_mask) == 0) atomicAdd(&s_base[digit], __popc(peers));
      }
      __syncthreads();
      for (int dg = tid; dg < RADIX; dg += SORT_THREADS) my_hist[dg] = s_base[dg];
    }
    if (SPLIT) {
      if (only_part == 0) return;
      __syncthreads();   // (s_wc is zeroed)
    } else {
      grid_barrier(barrier, ++epoch * gridDim.x);
    }

    // (c) my scatter bases: exclusive scan over digits of the problem-wide totals + lower CTAs' counts.
    // Thread tid owns the DPT consecutive digits tid*DPT .. tid*DPT+DPT-1.
    constexpr int DPT = RADIX / SORT_THREADS;
    uint32_t tot[DPT], mine[DPT];
    uint32_t tsum = 0;
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
      const int dg = tid * DPT + j;
      const uint32_t* h = hist_g + (size_t)prob * bpp * RADIX + dg;
      uint32_t t_ = 0, m_ = 0;
      for (int b = 0; b < bpp; ++b) {
        const uint32_t c = h[(size_t)b * RADIX];
        if (b < blk) m_ += c;
        t_ += c;
      }
      tot[j] = t_; mine[j] = m_; tsum += t_;
    }
    uint32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    uint32_t run0 = incl - tsum;
    for (int w = 0; w < warp; ++w) run0 += s_scan[w];
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
      s_base[tid * DPT + j] = run0 + mine[j];
      run0 += tot[j];
    }
    __syncthreads();

    // (d) stable scatter, tile by tile
    for (int64_t i0 = begin; i0 < end; i0 += SORT_THREADS) {
      const int64_t i = i0 + tid;
      const bool valid = i < end;
      const int32_t key = valid ? src_ids[i] : 0;
      const int32_t pos = valid ? (src_pos ? src_pos[i] : (int32_t)i) : 0;
      const uint32_t digit = valid ? (((uint32_t)key >> shift) & dmask) : invalid_digit;
      const uint32_t peers = __match_any_sync(0xffffffffu, digit);
      const uint32_t lrank = __popc(peers & lt_mask);
      const bool leader = valid && lrank == 0;
      if (leader) s_wc[warp][digit] = __popc(peers);
      __syncthreads();
      for (int dg = tid; dg < RADIX; dg += SORT_THREADS) {
        uint32_t run = s_base[dg];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
          const uint32_t c = s_wc[w][dg];
          if (c) { s_wc[w][dg] = run; run += c; }
        }
        s_base[dg] = run;
      }
      __syncthreads();
      if (valid) {
        const uint32_t d = s_wc[warp][digit] + lrank;
        dst_ids[d] = key;
        dst_pos[d] = pos;
      }
      __syncwarp();
      if (leader) s_wc[warp][digit] = 0;
    }
    if (!SPLIT) grid_barrier(barrier, ++epoch * gridDim.x);
  }
}

// tf.unique outputs from the sorted pairs (parity API; single CTA, not on the hot step).
__global__ void __launch_bounds__(1024) unique_first_occurrence_kernel(const int32_t* __restrict__ sid,
                                                                       const int32_t* __restrict__ spos, int64_t n,
                                                                       int32_t* __restrict__ uniq,
                                                                       int32_t* __restrict__ idx,
                                                                       int32_t* __restrict__ n_uniq_dev,
                                                                       int32_t* __restrict__ scratch) {
  __shared__ int32_t s_warp[32];
  __shared__ int32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t b = tid; b < n; b += 1024) scratch[b] = 0;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  // is_first[b] = 1 where b is the first occurrence of its id (= the position of a run head)
  for (int64_t k = tid; k < n; k += 1024)
    if (k == 0 || sid[k] != sid[k - 1]) scratch[spos[k]] = 1;
  __syncthreads();
  // exclusive scan over batch positions -> rank in first-occurrence order
  for (int64_t b0 = 0; b0 < n; b0 += 1024) {
    const int64_t b = b0 + tid;
    const int32_t x = b < n ? scratch[b] : 0;
    int32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int32_t off = s_carry;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (b < n) scratch[b] = off + incl - x;
    __syncthreads();
    if (tid == 1023) s_carry = off + incl;
    __syncthreads();
  }
  if (tid == 0) *n_uniq_dev = s_carry;
  // every run head labels its run
  for (int64_t k = tid; k < n; k += 1024) {
    if (k == 0 || sid[k] != sid[k - 1]) {
      const int32_t id = sid[k];
      const int32_t r = scratch[spos[k]];
      uniq[r] = id;
      for (int64_t j = k; j < n && sid[j] == id; ++j) idx[spos[j]] = r;
    }
  }
}

static int bits_for(int64_t max_id) {
  int bits = 1;
  while (bits < 32 && ((int64_t)1 << bits) < max_id) ++bits;
  return bits;
}

}  // namespace tfr

using namespace tfr;

// How many CTAs of the sort can be resident on the current device at once (cached per device).  The kernel's grid
// barrier needs EVERY CTA of the grid to become resident: the grid is sized against this, so a device with fewer SMs
// (or a build that uses more shared memory) shrinks the grid instead of hanging.  CTAs that have to wait for a
// neighbour kernel (the table pass) to leave room are fine: nothing that runs beside the sort waits on it.
__global__ void __launch_bounds__(64) sort_header_zero_kernel(uint32_t* header) { header[threadIdx.x] = 0u; }

static int sort_resident_capacity() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cached[dev] > 0) return cached[dev];
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dedup_sort_kernel<false>, SORT_THREADS, 0) != cudaSuccess)
    return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  // one CTA per SM at most: beside a persistent neighbour only that much room is certain
  cached[dev] = per_sm > 0 ? sms : 0;
  return cached[dev];
}

static int sort_bpp(int64_t n, int capacity) {
  int64_t bpp = (n + 1023) / 1024;
  if (bpp < 1) bpp = 1;
  if (bpp > SORT_MAX_BPP) bpp = SORT_MAX_BPP;
  if (bpp > capacity / 2) bpp = capacity / 2;
  return (int)bpp;
}

extern "C" int64_t tfr_dedup_workspace_bytes(int64_t n) {
  if (n < 0) return TFR_ERR_INVALID;
  return 256 + align_up(2 * SORT_MAX_BPP * RADIX * (int64_t)sizeof(uint32_t), 256) + 4 * align_up(n * 4, 256) + 256;
}

extern "C" int tfr_dedup_sort_pairs(const int32_t* ids_a, int64_t max_id_a, int32_t* sorted_ids_a,
                                    int32_t* sorted_pos_a, const int32_t* ids_b, int64_t max_id_b,
                                    int32_t* sorted_ids_b, int32_t* sorted_pos_b, int64_t n, void* workspace,
                                    int64_t workspace_bytes, void* stream) {
  return tfr_dedup_sort_pairs_tl(ids_a, max_id_a, sorted_ids_a, sorted_pos_a, ids_b, max_id_b, sorted_ids_b,
                                 sorted_pos_b, n, workspace, workspace_bytes, nullptr, stream);
}

extern "C" int tfr_dedup_sort_pairs_tl(const int32_t* ids_a, int64_t max_id_a, int32_t* sorted_ids_a,
                                       int32_t* sorted_pos_a, const int32_t* ids_b, int64_t max_id_b,
                                       int32_t* sorted_ids_b, int32_t* sorted_pos_b, int64_t n, void* workspace,
                                       int64_t workspace_bytes, const tfr_opt_scalars* opt, void* stream) {
  return dedup_sort_pairs_impl(ids_a, max_id_a, sorted_ids_a, sorted_pos_a, ids_b, max_id_b, sorted_ids_b, sorted_pos_b, n,
                               workspace, workspace_bytes, opt, stream, tune(TUNE_SORT_SPLIT) != 0);
}

int tfr::dedup_sort_pairs_impl(const int32_t* ids_a, int64_t max_id_a, int32_t* sorted_ids_a, int32_t* sorted_pos_a,
                               const int32_t* ids_b, int64_t max_id_b, int32_t* sorted_ids_b, int32_t* sorted_pos_b,
                               int64_t n, void* workspace, int64_t workspace_bytes, const tfr_opt_scalars* opt,
                               void* stream, bool split) {
  TFR_CHECK_ARG(n >= 0 && n < ((int64_t)1 << 31));
  if (n == 0) return TFR_OK;
  TFR_CHECK_ARG(ids_a && sorted_ids_a && sorted_pos_a && workspace && max_id_a > 0);
  TFR_CHECK_ARG(!ids_b || (sorted_ids_b && sorted_pos_b && max_id_b > 0));
  if (workspace_bytes < tfr_dedup_workspace_bytes(n)) {
    set_error("dedup workspace too small: %lld < %lld", (long long)workspace_bytes,
              (long long)tfr_dedup_workspace_bytes(n));
    return TFR_ERR_WORKSPACE;
  }
  char* w = reinterpret_cast<char*>(align_up((int64_t)(uintptr_t)workspace, 256));
  unsigned int* barrier = reinterpret_cast<unsigned int*>(w);
  w += 256;
  uint32_t* hist = reinterpret_cast<uint32_t*>(w);
  w += align_up(2 * SORT_MAX_BPP * RADIX * (int64_t)sizeof(uint32_t), 256);
  const int64_t seg = align_up(n * 4, 256);
  SortProblem pa{ids_a, sorted_ids_a, sorted_pos_a, (int32_t*)w, (int32_t*)(w + seg), n};
  SortProblem pb{ids_b, sorted_ids_b, sorted_pos_b, (int32_t*)(w + 2 * seg), (int32_t*)(w + 3 * seg), ids_b ? n : 0};
  int bits = bits_for(max_id_a);
  if (ids_b && bits_for(max_id_b) > bits) bits = bits_for(max_id_b);
  int n_passes = (bits + RADIX_BITS_MAX - 1) / RADIX_BITS_MAX;
  int digit_bits = (bits + n_passes - 1) / n_passes;
  const int capacity = sort_resident_capacity();
  if (capacity < 2) {
    set_error("id sort: the device cannot hold the sort's grid (occupancy query failed or < 2 resident CTAs)");
    return TFR_ERR_CUDA;
  }
  int bpp = sort_bpp(n, capacity);
  // the whole 256-byte header: the grid barrier and the fix-up work-list counters the segment sums keep at +64
  // (a 64-thread kernel of ours with the step kernels' carve-out: a plain kernel node in the captured graphs)
  TFR_PREP(sort_header_zero_kernel);
  sort_header_zero_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint32_t*>(barrier));
  TFR_LAUNCH_CHECK();
  if (split) {
    TFR_PREP(dedup_sort_kernel<true>);
    for (int pass = 0; pass < n_passes; ++pass)
      for (int part = 0; part < 2; ++part) {
        dedup_sort_kernel<true><<<2 * bpp, SORT_THREADS, 0, (cudaStream_t)stream>>>(pa, pb, bpp, n_passes, digit_bits, hist,
                                                                                    opt, barrier, pass, part);
        TFR_LAUNCH_CHECK();
      }
    return TFR_OK;
  }
  TFR_PREP(dedup_sort_kernel<false>);
  dedup_sort_kernel<false><<<2 * bpp, SORT_THREADS, 0, (cudaStream_t)stream>>>(pa, pb, bpp, n_passes, digit_bits, hist, opt,
                                                                               barrier, -1, -1);
  TFR_LAUNCH_CHECK();
  return TFR_OK;
}

extern "C" int tfr_unique_first_occurrence(const int32_t* sorted_ids, const int32_t* sorted_pos, int64_t n,
                                           int32_t* uniq, int32_t* idx, int32_t* n_uniq_dev, int32_t* scratch,
                                           void* stream) {
  TFR_CHECK_ARG(n >= 0 && n_uniq_dev);
  if (n == 0) {
    TFR_CUDA(cudaMemsetAsync(n_uniq_dev, 0, sizeof(int32_t), (cudaStream_t)stream));
    return TFR_OK;
  }
  TFR_CHECK_ARG(sorted_ids && sorted_pos && uniq && idx && scratch);
  unique_first_occurrence_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(sorted_ids, sorted_pos, n, uniq, idx, n_uniq_dev,
                                                                     scratch);
  TFR_LAUNCH_CHECK();
  return TFR_OK;
}
