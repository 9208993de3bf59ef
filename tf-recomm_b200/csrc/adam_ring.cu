// K4 (ring): the table-wide TF-Adam pass (TF: adam.py::_apply_sparse_shared) over INTERLEAVED feature
// tables [rows][var | m | v][width] as a TMA-bulk pipeline: one persistent CTA per SM, ONE thread moves the bytes --
// cp.async.bulk (1-D TMA, SASS UBLKCP) global -> shared into a ring of stages, mbarrier complete_tx signalling, and
// cp.async.bulk shared -> global for the updated stage -- while the consumer warps do nothing but LDS -> TF's
// arithmetic -> STS.  What this buys over the LDG/STG pass (adam.cu):
//   * bytes in flight are bounded by shared memory (n_ring x stage bytes, ~150 KB per SM), not by registers: the
//     LDG pass holds 6 x 16 B per thread and trip (86 KB per SM at 2 x 448 threads);
//   * every request is a multi-KB contiguous burst (a stage = 16 whole rows = 24 KB at dim 128), both directions;
//   * the slot-map lookup and the summed-gradient fetch of a touched row are prefetched one stage ahead, off the
//     arithmetic's critical path.
// Arithmetic, rounding order and the end-of-step work are the LDG pass's (adam_math.cuh): the two passes produce
// bit-identical tables (tests/test_gpu_parity.py::test_ring_pass_bit_identical).
// Bias tables (width 1, plain arrays) are swept by the consumer warps with 128-bit LDG/STG while the ring's first
// stages are in flight.
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>

#include "adam_math.cuh"

namespace tfr {

constexpr int RING_MAX_THREADS = 1024;
constexpr int RING_MAXU = 2;  // units (of 4 floats) per consumer thread and stage whose slot / gsum are prefetched

struct RingTable {
  float* base;          // var of row 0; row r: var at base + r*3*width, m at + width, v at + 2*width
  const int64_t* slot;  // [rows] (step stamp << 32 | run-head index into gsum); null = no row has a gradient
  const float* gsum;    // [n, width]
  int64_t rows;
  int64_t n_stages;     // ceil(rows / stage_rows)
  int32_t width;        // floats, % 4 == 0
  int32_t stage_rows;
};
struct RingBias {
  float *var, *m, *v;   // [rows]
  const int64_t* slot;
  const float* gsum;    // [n]
  int64_t rows;
};
struct RingArgs {
  RingTable t[2];
  RingBias b[2];
  int n_tabs, n_bias;
  int n_ring;            // stages in the ring
  uint32_t stage_bytes;  // bytes of one ring slot: [rows' var | m | v ... ][slot-map entries of the rows]
  uint32_t slot_off;     // offset of the slot-map entries inside a ring slot
  int l2_hint;           // L2 evict_first policy: bit 0 = on the bulk loads, bit 1 = on the bulk stores
  int slot_mode;         // 1 = slot-map entries ride along with the stage (default); experiments: 0 = read from global
                         // memory, 2 = as 1 without the look-ahead gradient fetch, 3 = no lookups at all (WRONG
                         // results: the memory ceiling of the pass)
  int copy_only;         // experiment: no arithmetic
  FinishArgs fin;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {  // non-blocking
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, bool hint,
                                          uint64_t pol) {
  if (hint)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
                     "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
  else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                     "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes, bool hint, uint64_t pol) {
  if (hint)
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes), "l"(pol) : "memory");
  else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory (the consumers' STS) -> visible to the async proxy (the bulk store)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// which table and which of its stages global stage g is, and where that stage lives
struct StageRef {
  int tb;
  int64_t row0;
  uint32_t rows;
};
__device__ __forceinline__ StageRef stage_ref(const RingArgs& a, int64_t g) {
  StageRef r;
  r.tb = (a.n_tabs > 1 && g >= a.t[0].n_stages) ? 1 : 0;
  const RingTable& t = a.t[r.tb];
  const int64_t ls = r.tb ? g - a.t[0].n_stages : g;
  r.row0 = ls * t.stage_rows;
  const int64_t left = t.rows - r.row0;
  r.rows = (uint32_t)(left < t.stage_rows ? left : t.stage_rows);
  return r;
}

// CAP / MINB: launch bounds only (register budget): <384, 2> = two CTAs of 8 consumer warps + the two DMA warps per SM,
// <576, 1> one CTA of up to 18 warps, <1024, 1> anything wider
template <int CAP, int MINB>
__global__ void __launch_bounds__(CAP, MINB) adam_ring_kernel(const __grid_constant__ RingArgs a,
                                                                       const tfr_opt_scalars* __restrict__ opt,
                                                                       int tl_slot) {
  pdl_wait();                // the fix-up's summed gradients and slot maps are complete
  pdl_launch_dependents();   // the next step's segment sums may become resident as this grid drains
  TlScope tl_scope(opt, tl_slot);
  extern __shared__ __align__(128) unsigned char ring_smem[];
  uint64_t* const full = reinterpret_cast<uint64_t*>(ring_smem + (size_t)a.n_ring * a.stage_bytes);
  uint64_t* const done = full + a.n_ring;
  uint64_t* const empty = done + a.n_ring;
  const int n_cons = (int)blockDim.x - 64;  // the last two warps move the bytes: one thread loads, one stores
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < a.n_ring; ++s) {
      mbar_init(full + s, 1);
      mbar_init(done + s, (uint32_t)(n_cons >> 5));
      mbar_init(empty + s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t total = a.t[0].n_stages + (a.n_tabs > 1 ? a.t[1].n_stages : 0);
  const int64_t n_it = total > (int64_t)blockIdx.x ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid >= n_cons) {
    auto gaddr = [&](const StageRef& r) {
      return reinterpret_cast<unsigned char*>(a.t[r.tb].base) + (size_t)r.row0 * (size_t)(12 * a.t[r.tb].width);
    };
    const uint64_t pol = a.l2_hint ? l2_evict_first_policy() : 0;
    if (tid == n_cons) {
      // ---- the LOADER: fills ring slot i % n_ring as soon as the storer has drained its previous contents ----
      const bool hl = a.l2_hint & 1;
      for (int64_t i = 0; i < n_it; ++i) {
        const int s = (int)(i % a.n_ring);
        if (i >= a.n_ring) mbar_wait(empty + s, (uint32_t)(((i / a.n_ring) - 1) & 1));
        const StageRef r = stage_ref(a, (int64_t)blockIdx.x + i * gridDim.x);
        const uint32_t bytes = r.rows * (uint32_t)(12 * a.t[r.tb].width);
        // the rows' slot-map entries ride along (an even number of them: 16-byte granularity; an odd last row's
        // entry is read from global memory by its consumers)
        const uint32_t sbytes = (a.slot_mode >= 1 && a.t[r.tb].slot) ? (r.rows & ~1u) * 8u : 0u;
        unsigned char* const dst = ring_smem + (size_t)s * a.stage_bytes;
        mbar_expect_tx(full + s, bytes + sbytes);
        bulk_load(dst, gaddr(r), bytes, full + s, hl, pol);
        if (sbytes) bulk_load(dst + a.slot_off, a.t[r.tb].slot + r.row0, sbytes, full + s, false, pol);
      }
    } else if (tid == n_cons + 32) {
      // ---- the STORER: writes a stage back once every consumer warp is done with it; the ring slot is handed
      // back to the loader when the bulk store has finished READING shared memory (one store group behind, so
      // that a store is always in flight) ----
      const bool hs = a.l2_hint & 2;
      for (int64_t i = 0; i < n_it; ++i) {
        const int s = (int)(i % a.n_ring);
        mbar_wait(done + s, (uint32_t)((i / a.n_ring) & 1));
        const StageRef r = stage_ref(a, (int64_t)blockIdx.x + i * gridDim.x);
        bulk_store(gaddr(r), ring_smem + (size_t)s * a.stage_bytes, r.rows * (uint32_t)(12 * a.t[r.tb].width), hs, pol);
        bulk_commit();
        if (i >= 1) {
          bulk_wait_read<1>();
          mbar_arrive(empty + (int)((i - 1) % a.n_ring));
        }
      }
      bulk_wait_all();
    }
  } else {
    const AdamK k = load_k(opt);
    const uint32_t stamp = (uint32_t)opt->global_step;
    // ---- bias tables first (the ring's first stages are in flight meanwhile): 4 rows per thread, 128-bit accesses ----
    {
      const int64_t gthreads = (int64_t)gridDim.x * n_cons;
      const int64_t gtid = (int64_t)blockIdx.x * n_cons + tid;
      for (int bt = 0; bt < a.n_bias; ++bt) {
        const RingBias& b = a.b[bt];
        const bool vec_ok = ((((uintptr_t)b.var) | ((uintptr_t)b.m) | ((uintptr_t)b.v)) & 15u) == 0;
        const int64_t n4 = vec_ok ? (b.rows >> 2) : 0;
        for (int64_t q = gtid; q < n4; q += gthreads) {
          float4 x = ld_stream_f4(reinterpret_cast<const float4*>(b.var) + q);
          float4 y = ld_stream_f4(reinterpret_cast<const float4*>(b.m) + q);
          float4 z = ld_stream_f4(reinterpret_cast<const float4*>(b.v) + q);
          int64_t sl[4] = {-1, -1, -1, -1};
          if (b.slot) {
#pragma unroll
            for (int c = 0; c < 4; ++c) sl[c] = b.slot[4 * q + c];
          }
          float* xs[4] = {&x.x, &x.y, &x.z, &x.w};
          float* ys[4] = {&y.x, &y.y, &y.z, &y.w};
          float* zs[4] = {&z.x, &z.y, &z.z, &z.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (b.slot && (uint32_t)(sl[c] >> 32) == stamp) adam_grad(*xs[c], *ys[c], *zs[c], b.gsum[(uint32_t)sl[c]], k);
            else adam_decay(*xs[c], *ys[c], *zs[c], k);
          }
          st_stream_f4(reinterpret_cast<float4*>(b.var) + q, x);
          st_stream_f4(reinterpret_cast<float4*>(b.m) + q, y);
          st_stream_f4(reinterpret_cast<float4*>(b.v) + q, z);
        }
        for (int64_t e = 4 * n4 + gtid; e < b.rows; e += gthreads) {  // the rows % 4 tail (or everything, unaligned)
          float p = b.var[e], q = b.m[e], r = b.v[e];
          const int64_t sl = b.slot ? b.slot[e] : -1;
          if (b.slot && (uint32_t)(sl >> 32) == stamp) adam_grad(p, q, r, b.gsum[(uint32_t)sl], k);
          else adam_decay(p, q, r, k);
          b.var[e] = p; b.m[e] = q; b.v[e] = r;
        }
      }
    }
    // ---- the ring's consumers: unit u of a stage (row u / upr, column u % upr) belongs to thread u % n_cons ----
    // A row's slot-map entry arrives in shared memory with the stage.  The summed gradient of a touched row is
    // fetched one stage AHEAD when that stage has already landed (the loads run n_ring - 1 stages ahead, so it
    // usually has): the L2 round trip then overlaps this stage's arithmetic instead of preceding the next one's.
    float4 g_pf[RING_MAXU];
    bool has_pf[RING_MAXU];
    bool pf_valid = false;
#pragma unroll
    for (int j = 0; j < RING_MAXU; ++j) { has_pf[j] = false; g_pf[j] = make_float4(0.f, 0.f, 0.f, 0.f); }
    // slot-map entry of row `row` of the stage in ring slot `st_base`
    auto slot_of = [&](const RingTable& t, const StageRef& r, const unsigned char* st_base, uint32_t row) -> int64_t {
      if (!t.slot || a.slot_mode == 3) return -1;
      if (a.slot_mode >= 1 && row < (r.rows & ~1u)) return reinterpret_cast<const int64_t*>(st_base + a.slot_off)[row];
      return t.slot[r.row0 + row];
    };
    auto lookup = [&](int64_t i, bool (&has)[RING_MAXU], float4 (&g)[RING_MAXU]) {
      const StageRef r = stage_ref(a, (int64_t)blockIdx.x + i * gridDim.x);
      const RingTable& t = a.t[r.tb];
      const uint32_t upr = (uint32_t)t.width >> 2, units = r.rows * upr;
      const unsigned char* st_base = ring_smem + (size_t)(i % a.n_ring) * a.stage_bytes;
#pragma unroll
      for (int j = 0; j < RING_MAXU; ++j) {
        has[j] = false;
        const uint32_t u = (uint32_t)tid + (uint32_t)j * (uint32_t)n_cons;
        if (u < units) {
          const uint32_t row = u / upr;
          const int64_t sl = slot_of(t, r, st_base, row);
          if ((uint32_t)(sl >> 32) == stamp && t.slot) {
            has[j] = true;
            g[j] = *reinterpret_cast<const float4*>(t.gsum + (size_t)(uint32_t)sl * t.width + ((u - row * upr) << 2));
          }
        }
      }
    };
    for (int64_t i = 0; i < n_it; ++i) {
      const StageRef r = stage_ref(a, (int64_t)blockIdx.x + i * gridDim.x);
      const RingTable& t = a.t[r.tb];
      const uint32_t upr = (uint32_t)t.width >> 2, units = r.rows * upr;
      const int s = (int)(i % a.n_ring);
      unsigned char* const st_base = ring_smem + (size_t)s * a.stage_bytes;
      float4* const st = reinterpret_cast<float4*>(st_base);
      mbar_wait(full + s, (uint32_t)((i / a.n_ring) & 1));
      float4 g[RING_MAXU];
      bool has[RING_MAXU];
      if (pf_valid) {
#pragma unroll
        for (int j = 0; j < RING_MAXU; ++j) { has[j] = has_pf[j]; g[j] = g_pf[j]; }
      } else {
        lookup(i, has, g);
      }
      pf_valid = false;
      if (a.slot_mode <= 1 && i + 1 < n_it && mbar_test(full + (int)((i + 1) % a.n_ring), (uint32_t)(((i + 1) / a.n_ring) & 1))) {
        lookup(i + 1, has_pf, g_pf);
        pf_valid = true;
      }
      auto update = [&](uint32_t u, bool hs, const float4& gg) {
        const uint32_t row = u / upr, col = u - row * upr;
        float4* const p = st + (size_t)row * 3u * upr + col;
        float4 x = p[0], y = p[upr], z = p[2u * upr];
        if (a.copy_only) {
        } else if (hs) {
          adam_grad(x.x, y.x, z.x, gg.x, k);
          adam_grad(x.y, y.y, z.y, gg.y, k);
          adam_grad(x.z, y.z, z.z, gg.z, k);
          adam_grad(x.w, y.w, z.w, gg.w, k);
        } else {
          adam_decay(x.x, y.x, z.x, k);
          adam_decay(x.y, y.y, z.y, k);
          adam_decay(x.z, y.z, z.z, k);
          adam_decay(x.w, y.w, z.w, k);
        }
        p[0] = x; p[upr] = y; p[2u * upr] = z;
      };
#pragma unroll
      for (int j = 0; j < RING_MAXU; ++j) {
        const uint32_t u = (uint32_t)tid + (uint32_t)j * (uint32_t)n_cons;
        if (u < units) update(u, has[j], g[j]);
      }
      // more units per thread than prefetch registers (a wide stage with a narrow CTA): look the row up here
      for (uint32_t u = (uint32_t)tid + (uint32_t)RING_MAXU * (uint32_t)n_cons; u < units; u += (uint32_t)n_cons) {
        bool hs = false;
        float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t row = u / upr;
        const int64_t s2 = slot_of(t, r, st_base, row);
        if ((uint32_t)(s2 >> 32) == stamp && t.slot) {
          hs = true;
          gg = *reinterpret_cast<const float4*>(t.gsum + (size_t)(uint32_t)s2 * t.width + ((u - row * upr) << 2));
        }
        update(u, hs, gg);
      }
      fence_async_smem();
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(done + s);
    }
  }
  if (a.fin.n_partials > 0) {  // the last CTA to get here ends the step (every CTA has read lr_t and the stamp by now)
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      s_last = atomicAdd(&a.fin.opt->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
      finish_step_scalars(a.fin.mu, a.fin.m_mu, a.fin.v_mu, a.fin.opt, a.fin.partials, a.fin.se_partials,
                          a.fin.n_partials);
      if (threadIdx.x == 0) a.fin.opt->ticket = 0;
    }
  }
}

struct RingConfig {
  int enabled, n_ring, stage_kb, threads, l2_hint, ctas_per_sm, slot_mode, copy_only;
};
static RingConfig ring_config() {
  RingConfig c;
  c.enabled = tune(TUNE_PASS_RING);
  c.n_ring = tune(TUNE_RING_STAGES);
  c.stage_kb = tune(TUNE_RING_STAGE_KB);
  c.threads = tune(TUNE_RING_THREADS);   // consumer warps + the loader warp + the storer warp
  c.l2_hint = tune(TUNE_RING_L2_HINT);
  c.ctas_per_sm = tune(TUNE_RING_CTAS_PER_SM);
  c.slot_mode = tune(TUNE_RING_SLOT_MODE);
  c.copy_only = tune(TUNE_STREAM_COPY_ONLY);
  return c;
}

// Can this set of tables go through the ring pass?  Interleaved feature tables (m = var + width, v = var + 2*width,
// stride 3*width, width % 4 == 0, 16-byte aligned) + plain bias tables.
bool ring_pass_eligible(const tfr_adam_table* tabs, int n) {
  if (!ring_config().enabled) return false;
  int nf = 0, nb = 0;
  for (int i = 0; i < n; ++i) {
    const tfr_adam_table& t = tabs[i];
    if (t.rows == 0) continue;
    if (t.width == 1 && (t.stride == 0 || t.stride == 1)) { if (++nb > 2) return false; continue; }
    if (t.width % 4 != 0 || t.stride != 3 * (int64_t)t.width || t.m != t.var + t.width || t.v != t.var + 2 * t.width ||
        ((uintptr_t)t.var & 15u) || (t.gsum && ((uintptr_t)t.gsum & 15u)))
      return false;
    if (2 * 12 * (int64_t)t.width > 48 * 1024) return false;  // a stage is at least two rows
    if (t.slot && ((uintptr_t)t.slot & 15u)) return false;
    if (++nf > 2) return false;
  }
  return nf > 0;
}

int ring_pass_launch(const tfr_adam_table* tabs, int n, const tfr_opt_scalars* opt, int tl_slot, cudaStream_t st,
                     const FinishArgs* fin) {
  const RingConfig c = ring_config();
  RingArgs a;
  memset(&a, 0, sizeof(a));
  if (fin) a.fin = *fin;
  uint32_t stage_bytes = 0;
  int max_stage_rows = 0;
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    const tfr_adam_table& t = tabs[i];
    if (t.rows == 0) continue;
    if (t.width == 1) {
      a.b[a.n_bias++] = RingBias{t.var, t.m, t.v, t.slot, t.gsum, t.rows};
      continue;
    }
    RingTable& r = a.t[a.n_tabs++];
    r.base = t.var; r.slot = t.slot; r.gsum = t.gsum; r.rows = t.rows; r.width = t.width;
    const int row_bytes = 12 * t.width;
    int sr = (c.stage_kb * 1024) / row_bytes;
    sr &= ~1;  // even: the slot-map entries of a stage are copied in 16-byte units
    if (sr < 2) sr = 2;
    r.stage_rows = sr;
    if (sr > max_stage_rows) max_stage_rows = sr;
    r.n_stages = (t.rows + sr - 1) / sr;
    total += r.n_stages;
    if ((uint32_t)(sr * row_bytes) > stage_bytes) stage_bytes = (uint32_t)(sr * row_bytes);
  }
  stage_bytes = (stage_bytes + 127u) & ~127u;
  a.slot_off = stage_bytes;
  stage_bytes += ((uint32_t)max_stage_rows * 8u + 127u) & ~127u;
  a.stage_bytes = stage_bytes;
  a.n_ring = c.n_ring < 2 ? 2 : c.n_ring;
  a.l2_hint = c.l2_hint;
  a.slot_mode = c.slot_mode;
  a.copy_only = c.copy_only;
  const size_t smem = (size_t)a.n_ring * stage_bytes + 3 * (size_t)a.n_ring * sizeof(uint64_t);
  if (smem > 227 * 1024) {
    set_error("ring pass: %zu bytes of shared memory (RING_STAGES x RING_STAGE_KB too large)", smem);
    return TFR_ERR_INVALID;
  }
  int64_t grid = (int64_t)sm_count() * (c.ctas_per_sm < 1 ? 1 : c.ctas_per_sm);
  if (grid > total) grid = total;
  if (grid < 1) grid = 1;
  int threads = c.threads;
  if (threads < 96) threads = 96;
  if (threads > RING_MAX_THREADS) threads = RING_MAX_THREADS;
  threads = threads / 32 * 32;
  static std::mutex mu;
  static std::map<const void*, size_t> prepared;
  auto go = [&](auto kernel) -> int {
    const void* fn = reinterpret_cast<const void*>(kernel);
    {
      std::lock_guard<std::mutex> g(mu);
      if (smem > prepared[fn]) {
        TFR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prepared[fn] = smem;
      }
    }
    prep_kernel_carveout(fn, smem * (size_t)(c.ctas_per_sm < 1 ? 1 : c.ctas_per_sm));
    TFR_CUDA(launch_kernel(kernel, dim3((unsigned)grid), dim3(threads), smem, st, tune(TUNE_PDL) != 0, a, opt, tl_slot));
    return TFR_OK;
  };
  if (threads <= 320 && c.ctas_per_sm >= 2) return go(adam_ring_kernel<384, 2>);
  if (threads <= 576) return go(adam_ring_kernel<576, 1>);
  return go(adam_ring_kernel<1024, 1>);
}

}  // namespace tfr
