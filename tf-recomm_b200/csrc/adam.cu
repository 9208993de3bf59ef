// K4: TF AdamOptimizer's IndexedSlices semantics (TF: adam.py::_apply_sparse_shared; the
// optimizer north_star names, selected at ops.py:144/148):
//     m   = m*beta1                (WHOLE table)      m[uq]  += g*(1-beta1)
//     v   = v*beta2                (WHOLE table)      v[uq]  += (g*g)*(1-beta2)
//     var = var - (lr_t*m)/(sqrt(v)+eps)  (WHOLE table)
// TF runs this as ~8 unfused table-wide kernels; here each parameter is read and written exactly once
// per step (24 B/param): adam_stream_kernel is the streaming pass over the rows NOT in this step's
// slice (g = 0), adam_touched_kernel updates the slice rows with their summed gradient.  Every
// operation keeps TF's rounding order (separate mul/add, correctly rounded sqrt and divide).
// Also: dense ApplyAdam for bias_global (TF: training_ops.cc, A.5), scatter_sub SGD (ops.py:145) and
// the end-of-step bookkeeping (TF: adam.py::_finish).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace tfr {

struct AdamK {
  float b1, b2, lr_t, eps, omb1, omb2;
};
__device__ __forceinline__ AdamK load_k(const tfr_opt_scalars* opt) {
  AdamK k;
  k.b1 = opt->beta1; k.b2 = opt->beta2; k.lr_t = opt->lr_t; k.eps = opt->eps;
  k.omb1 = opt->one_minus_beta1; k.omb2 = opt->one_minus_beta2;
  return k;
}
__device__ __forceinline__ void adam_decay(float& var, float& m, float& v, const AdamK& k) {
  m = mul_rn(m, k.b1);
  v = mul_rn(v, k.b2);
  var = sub_rn(var, div_rn(mul_rn(k.lr_t, m), add_rn(sqrt_rn(v), k.eps)));
}
__device__ __forceinline__ void adam_grad(float& var, float& m, float& v, float g, const AdamK& k) {
  m = add_rn(mul_rn(m, k.b1), mul_rn(g, k.omb1));
  v = add_rn(mul_rn(v, k.b2), mul_rn(mul_rn(g, g), k.omb2));
  var = sub_rn(var, div_rn(mul_rn(k.lr_t, m), add_rn(sqrt_rn(v), k.eps)));
}

// ---- the whole-table pass: every row of every table, in address order, ONE launch ----------------------
// Each parameter is read and written exactly once per step (24 B/param).  The concatenated tables are cut
// into units of 4 floats; a persistent grid strides over them, UNROLL units per thread and trip (6*UNROLL
// 16-byte requests in flight per thread), .cs (evict-first) accesses because nothing is reused.  A row of
// this step's slice (slot[row] = k >= 0) adds its summed gradient gsum[k]; all other rows get g = 0, for
// which m*b1 + 0*(1-b1), v*b2 + 0 reduce to TF's pure decay bit for bit.  Splitting the slice rows into a
// separate scattered kernel was measured slower: both halves lose DRAM page locality (4.9 and 3.7 TB/s
// against 6.0 TB/s for the in-order pass).
struct StreamTab {
  float *var, *m, *v;
  const int32_t* slot;
  const float* gsum;
  uint32_t n;        // floats
  uint32_t width;    // floats per row
  uint32_t unit_end; // exclusive end of this table's units in the concatenated unit space
};
struct StreamArgs {
  StreamTab t[4];
  int n_tabs;
  uint32_t total_units;
};

template <int UNROLL>
__global__ void __launch_bounds__(512, 2) adam_stream_multi_kernel(const __grid_constant__ StreamArgs a,
                                                                const tfr_opt_scalars* __restrict__ opt, int tl_slot) {
  TlScope tl_scope(opt, tl_slot);
  const AdamK k = load_k(opt);
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t q0 = blockIdx.x * blockDim.x + threadIdx.x; q0 < a.total_units; q0 += stride * UNROLL) {
    float4 x[UNROLL], y[UNROLL], z[UNROLL], g[UNROLL];
    uint32_t lu[UNROLL];
    int tab[UNROLL];
    bool full[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint32_t q = q0 + u * stride;
      tab[u] = -1;
      if (q >= a.total_units) continue;
      int tb = 0;
      uint32_t begin = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (tb == i && i + 1 < a.n_tabs && q >= a.t[i].unit_end) { begin = a.t[i].unit_end; tb = i + 1; }
      tab[u] = tb;
      lu[u] = q - begin;
      const StreamTab& t = a.t[tb];
      full[u] = lu[u] * 4u + 3u < t.n;
      if (full[u]) {
        x[u] = ld_stream_f4(reinterpret_cast<const float4*>(t.var) + lu[u]);
        y[u] = ld_stream_f4(reinterpret_cast<const float4*>(t.m) + lu[u]);
        z[u] = ld_stream_f4(reinterpret_cast<const float4*>(t.v) + lu[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      g[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (tab[u] < 0) continue;
      const StreamTab& t = a.t[tab[u]];
      if (!t.slot) continue;
      const uint32_t e0 = lu[u] * 4u;
      if ((t.width & 3u) == 0u) {  // the unit lies in one row
        const uint32_t row = e0 / t.width;
        const int32_t sl = t.slot[row];
        if (sl >= 0) g[u] = *reinterpret_cast<const float4*>(t.gsum + (size_t)sl * t.width + (e0 - row * t.width));
      } else {                     // dim 15, bias tables: every float has its own row
        float gg[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t e = e0 + j;
          if (e < t.n) {
            const uint32_t row = e / t.width;
            const int32_t sl = t.slot[row];
            if (sl >= 0) gg[j] = t.gsum[(size_t)sl * t.width + (e - row * t.width)];
          }
        }
        g[u] = make_float4(gg[0], gg[1], gg[2], gg[3]);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (tab[u] < 0) continue;
      const StreamTab& t = a.t[tab[u]];
      if (full[u]) {
        adam_grad(x[u].x, y[u].x, z[u].x, g[u].x, k);
        adam_grad(x[u].y, y[u].y, z[u].y, g[u].y, k);
        adam_grad(x[u].z, y[u].z, z[u].z, g[u].z, k);
        adam_grad(x[u].w, y[u].w, z[u].w, g[u].w, k);
        st_stream_f4(reinterpret_cast<float4*>(t.var) + lu[u], x[u]);
        st_stream_f4(reinterpret_cast<float4*>(t.m) + lu[u], y[u]);
        st_stream_f4(reinterpret_cast<float4*>(t.v) + lu[u], z[u]);
      } else {  // the table's last, partial unit
        const float gg[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t e = lu[u] * 4u + j;
          if (e < t.n) {
            float p = t.var[e], q = t.m[e], r = t.v[e];
            adam_grad(p, q, r, gg[j], k);
            t.var[e] = p; t.m[e] = q; t.v[e] = r;
          }
        }
      }
    }
  }
}

// ---- slice rows: one lane group per ENT consecutive sorted entries; run heads act --------------------------
// Both tables' slices (users, items) in ONE launch (blockIdx.y).  A group first decides which of its ENT
// entries are run heads, then issues the loads of all of them together (up to 4*ENT 16-byte requests per
// lane in flight) before any arithmetic: the kernel is a chain of dependent latencies otherwise.
struct SliceSide {
  float *var, *m, *v;        // feature table (null = not trained)
  float *bvar, *bm, *bv;     // bias table sharing the row ids (null = not trained)
  const int32_t* sid;        // sorted ids
  const float* gsum;         // [n, width] summed gradient at run heads
  const float* bgsum;        // [n]
};
constexpr int SLICE_ENT = 4;

template <int VEC, int L>
__global__ void __launch_bounds__(256) adam_slice_kernel(SliceSide s0, SliceSide s1, int width, int64_t n,
                                                         const tfr_opt_scalars* __restrict__ opt, int sgd, int tl_slot) {
  TlScope tl_scope(opt, tl_slot);
  const SliceSide s = blockIdx.y ? s1 : s0;
  const int lane = threadIdx.x & (L - 1);
  const int64_t kk0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L) * SLICE_ENT;
  if (kk0 >= n) return;
  int32_t id[SLICE_ENT];
  bool head[SLICE_ENT];
  int32_t prev = kk0 > 0 ? s.sid[kk0 - 1] : -1;
#pragma unroll
  for (int e = 0; e < SLICE_ENT; ++e) {
    const int64_t kk = kk0 + e;
    id[e] = kk < n ? s.sid[kk] : -1;
    head[e] = kk < n && id[e] != prev;
    prev = id[e];
  }
  AdamK k;
  if (!sgd) k = load_k(opt);
  if (s.bvar && lane == 0) {  // the rows' bias entries (a width-1 table sharing the row ids)
    float a[SLICE_ENT], b[SLICE_ENT], d[SLICE_ENT], g[SLICE_ENT];
#pragma unroll
    for (int e = 0; e < SLICE_ENT; ++e)
      if (head[e]) {
        a[e] = s.bvar[id[e]];
        g[e] = s.bgsum[kk0 + e];
        if (!sgd) { b[e] = s.bm[id[e]]; d[e] = s.bv[id[e]]; }
      }
#pragma unroll
    for (int e = 0; e < SLICE_ENT; ++e)
      if (head[e]) {
        if (sgd) {
          s.bvar[id[e]] = sub_rn(a[e], g[e]);
        } else {
          adam_grad(a[e], b[e], d[e], g[e], k);
          s.bvar[id[e]] = a[e]; s.bm[id[e]] = b[e]; s.bv[id[e]] = d[e];
        }
      }
  }
  if (!s.var) return;
  const int n_units = width / VEC;
  for (int unit = lane; unit < n_units; unit += L) {
    float a[SLICE_ENT][VEC], b[SLICE_ENT][VEC], d[SLICE_ENT][VEC], g[SLICE_ENT][VEC];
#pragma unroll
    for (int e = 0; e < SLICE_ENT; ++e) {
      if (!head[e]) continue;
      const size_t off = (size_t)id[e] * width + (size_t)unit * VEC;
      const size_t goff = (size_t)(kk0 + e) * width + (size_t)unit * VEC;
      if constexpr (VEC == 4) {
        const float4 av = *reinterpret_cast<const float4*>(s.var + off);
        const float4 gv = *reinterpret_cast<const float4*>(s.gsum + goff);
        a[e][0] = av.x; a[e][1] = av.y; a[e][2] = av.z; a[e][3] = av.w;
        g[e][0] = gv.x; g[e][1] = gv.y; g[e][2] = gv.z; g[e][3] = gv.w;
        if (!sgd) {
          const float4 bv4 = *reinterpret_cast<const float4*>(s.m + off);
          const float4 dv4 = *reinterpret_cast<const float4*>(s.v + off);
          b[e][0] = bv4.x; b[e][1] = bv4.y; b[e][2] = bv4.z; b[e][3] = bv4.w;
          d[e][0] = dv4.x; d[e][1] = dv4.y; d[e][2] = dv4.z; d[e][3] = dv4.w;
        }
      } else {
        a[e][0] = s.var[off]; g[e][0] = s.gsum[goff];
        if (!sgd) { b[e][0] = s.m[off]; d[e][0] = s.v[off]; }
      }
    }
#pragma unroll
    for (int e = 0; e < SLICE_ENT; ++e) {
      if (!head[e]) continue;
      const size_t off = (size_t)id[e] * width + (size_t)unit * VEC;
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        if (sgd) a[e][c] = sub_rn(a[e][c], g[e][c]);  // ops.py:145 scatter_sub; gsum holds the sum of lr*g
        else adam_grad(a[e][c], b[e][c], d[e][c], g[e][c], k);
      }
      if constexpr (VEC == 4) {
        *reinterpret_cast<float4*>(s.var + off) = make_float4(a[e][0], a[e][1], a[e][2], a[e][3]);
        if (!sgd) {
          *reinterpret_cast<float4*>(s.m + off) = make_float4(b[e][0], b[e][1], b[e][2], b[e][3]);
          *reinterpret_cast<float4*>(s.v + off) = make_float4(d[e][0], d[e][1], d[e][2], d[e][3]);
        }
      } else {
        s.var[off] = a[e][0];
        if (!sgd) { s.m[off] = b[e][0]; s.v[off] = d[e][0]; }
      }
    }
  }
}

// ---- end of step ----------------------------------------------------------------------------------
// every CTA resets the slot-map entries of its part of the batch to -1; CTA 0 / warp 0 folds the per-CTA
// partials in a fixed order, applies the dense update of bias_global and advances the step scalars.
__global__ void __launch_bounds__(256) finish_step_kernel(tfr_svd_tables t, tfr_opt_scalars* opt,
                                                          const int32_t* __restrict__ users,
                                                          const int32_t* __restrict__ items, int64_t B,
                                                          const float* __restrict__ partials,
                                                          const double* __restrict__ se_partials, int n_partials) {
  TlScope tl_scope(opt, TFR_TL_FINISH);
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {  // ids >= the table size mark occurrences owned by another rank (row-sharded mode)
    const int32_t u = users[b], i = items[b];
    if (u < t.user_num) t.user_slot[u] = -1;
    if (i < t.item_num) t.item_slot[i] = -1;
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    float a = 0.0f;
    double se = 0.0;
    for (int j = threadIdx.x; j < n_partials; j += 32) { a = add_rn(a, partials[j]); se += se_partials[j]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = add_rn(a, __shfl_xor_sync(0xffffffffu, a, o));
      se += __shfl_xor_sync(0xffffffffu, se, o);
    }
    if (threadIdx.x == 0) {
      const float g = a;  // d cost / d bias_global = sum_b e_b  (A.3)
      opt->g_mu = g;
      opt->se_sum = se;
      if (opt->se_ring && opt->se_ring_len > 0) opt->se_ring[opt->global_step % opt->se_ring_len] = se;
      const bool sgd = opt->flags & TFR_OPT_SGD;
      if (opt->var_mask & TFR_VAR_MU) {
        if (sgd) {
          *t.mu = sub_rn(*t.mu, mul_rn(opt->lr, g));
        } else {  // TF: training_ops.cc ApplyAdam (A.5)
          float alpha = sqrt_rn(sub_rn(1.0f, opt->beta2_power));
          alpha = mul_rn(opt->lr, alpha);
          alpha = div_rn(alpha, sub_rn(1.0f, opt->beta1_power));
          float mm = *t.m_mu, vv = *t.v_mu;
          mm = add_rn(mm, mul_rn(sub_rn(g, mm), opt->one_minus_beta1));
          vv = add_rn(vv, mul_rn(sub_rn(mul_rn(g, g), vv), opt->one_minus_beta2));
          *t.m_mu = mm;
          *t.v_mu = vv;
          *t.mu = sub_rn(*t.mu, div_rn(mul_rn(mm, alpha), add_rn(sqrt_rn(vv), opt->eps)));
        }
      }
      if (!sgd) {  // TF: adam.py::_finish
        opt->beta1_power = mul_rn(opt->beta1_power, opt->beta1);
        opt->beta2_power = mul_rn(opt->beta2_power, opt->beta2);
        // lr_t of the NEXT step (TF: _apply_sparse_shared recomputes it from the advanced powers), so that a step
        // needs no kernel in front of the forward
        float tt = sqrt_rn(sub_rn(1.0f, opt->beta2_power));
        opt->lr_t = div_rn(mul_rn(opt->lr, tt), sub_rn(1.0f, opt->beta1_power));
      }
      opt->global_step += 1;
      opt->batch_cursor += 1;
    }
  }
}

__global__ void opt_init_kernel(tfr_opt_scalars* opt, float lr, float reg, float beta1, float beta2, float eps,
                                int flags, int var_mask) {
  opt->lr = lr; opt->reg = reg; opt->beta1 = beta1; opt->beta2 = beta2; opt->eps = eps;
  opt->beta1_power = beta1; opt->beta2_power = beta2;  // TF: adam.py _create_slots
  opt->one_minus_beta1 = sub_rn(1.0f, beta1);
  opt->one_minus_beta2 = sub_rn(1.0f, beta2);
  float tt = sqrt_rn(sub_rn(1.0f, beta2));
  opt->lr_t = div_rn(mul_rn(lr, tt), sub_rn(1.0f, beta1));
  opt->flags = flags; opt->var_mask = var_mask;
  opt->global_step = 0; opt->batch_cursor = 0;
  opt->se_sum = 0.0; opt->g_mu = 0.0f; opt->pad_ = 0.0f;
  opt->se_ring = nullptr; opt->se_ring_len = 0; opt->timeline = nullptr;
}

}  // namespace tfr

using namespace tfr;

extern "C" int tfr_opt_init(tfr_opt_scalars* opt_dev, float lr, float reg, float beta1, float beta2, float eps,
                            int32_t flags, int32_t var_mask, void* stream) {
  TFR_CHECK_ARG(opt_dev);
  opt_init_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(opt_dev, lr, reg, beta1, beta2, eps, flags, var_mask);
  TFR_LAUNCH_CHECK();
  return TFR_OK;
}

static int launch_stream_chunks(const StreamTab* chunks, int n_chunks, const tfr_opt_scalars* opt, int tl_slot,
                                cudaStream_t st) {
  StreamArgs a;
  memset(&a, 0, sizeof(a));
  uint64_t units = 0;
  for (int i = 0; i < n_chunks; ++i) {
    a.t[i] = chunks[i];
    units += ((uint64_t)chunks[i].n + 3) / 4;
    a.t[i].unit_end = (uint32_t)units;
  }
  a.n_tabs = n_chunks;
  a.total_units = (uint32_t)units;
  // persistent grid: 2 CTAs x 512 threads x 64 registers per SM, 2 units per thread and trip: 5.9 TB/s in situ.
  // (3 x 256 leaves room for another kernel's CTAs beside it but measures 7 % slower: TFR_STREAM_* to experiment.)
  static int cfg_ctas = -1, cfg_unroll = 0, cfg_threads = 512;
  if (cfg_ctas < 0) {
    const char* e1 = getenv("TFR_STREAM_CTAS_PER_SM");
    const char* e2 = getenv("TFR_STREAM_UNROLL");
    const char* e3 = getenv("TFR_STREAM_THREADS");
    cfg_ctas = e1 ? atoi(e1) : 2;
    cfg_unroll = e2 ? atoi(e2) : 2;
    cfg_threads = e3 ? atoi(e3) : 512;
  }
  int64_t grid = ((int64_t)units + cfg_threads * cfg_unroll - 1) / (cfg_threads * cfg_unroll);
  const int64_t cap = (int64_t)sm_count() * cfg_ctas;
  if (cfg_ctas > 0 && grid > cap) grid = cap;
  if (cfg_unroll >= 4) {
    TFR_PREP(adam_stream_multi_kernel<4>);
    adam_stream_multi_kernel<4><<<(unsigned)grid, cfg_threads, 0, st>>>(a, opt, tl_slot);
  } else {
    TFR_PREP(adam_stream_multi_kernel<2>);
    adam_stream_multi_kernel<2><<<(unsigned)grid, cfg_threads, 0, st>>>(a, opt, tl_slot);
  }
  TFR_LAUNCH_CHECK();
  return TFR_OK;
}

extern "C" int tfr_adam_stream_multi(const tfr_adam_table* tables, int32_t n_tables, const tfr_opt_scalars* opt,
                                     int32_t tl_slot, void* stream) {
  TFR_CHECK_ARG(tables && n_tables >= 1 && n_tables <= 4 && opt && tl_slot >= 0 && tl_slot < TFR_TL_SLOTS);
  // The kernel indexes floats with 32 bits: a table of >= 2^32 floats (the 50M x 128 user shard of configs[4] at
  // G = 2) is cut into row ranges, and launches are split so that one launch covers < 2^32 units.
  const uint64_t kMaxFloats = ((uint64_t)1 << 32) - 8;
  StreamTab pending[4];
  int np = 0;
  uint64_t pending_units = 0;
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < n_tables; ++i) {
    const tfr_adam_table& t = tables[i];
    TFR_CHECK_ARG(t.rows >= 0 && t.width > 0);
    if (t.rows == 0) continue;
    TFR_CHECK_ARG(t.var && t.m && t.v && (!t.slot || t.gsum));
    TFR_CHECK_ARG(((uintptr_t)t.var % 16 == 0) && ((uintptr_t)t.m % 16 == 0) && ((uintptr_t)t.v % 16 == 0));
    TFR_CHECK_ARG(!t.gsum || t.width % 4 != 0 || (uintptr_t)t.gsum % 16 == 0);
    uint64_t rows_per_chunk = kMaxFloats / (uint64_t)t.width;
    rows_per_chunk -= rows_per_chunk % 4;  // keeps every chunk's first float 16-byte aligned for any width
    for (uint64_t r0 = 0; r0 < (uint64_t)t.rows; r0 += rows_per_chunk) {
      const uint64_t nr = ((uint64_t)t.rows - r0 < rows_per_chunk) ? (uint64_t)t.rows - r0 : rows_per_chunk;
      StreamTab c;
      const uint64_t off = r0 * (uint64_t)t.width;
      c.var = t.var + off; c.m = t.m + off; c.v = t.v + off;
      c.slot = t.slot ? t.slot + r0 : nullptr;
      c.gsum = t.gsum;
      c.n = (uint32_t)(nr * (uint64_t)t.width);
      c.width = (uint32_t)t.width;
      c.unit_end = 0;
      const uint64_t cu = ((uint64_t)c.n + 3) / 4;
      if (np == 4 || pending_units + cu >= ((uint64_t)1 << 32)) {
        int rc = launch_stream_chunks(pending, np, opt, tl_slot, st);
        if (rc) return rc;
        np = 0;
        pending_units = 0;
      }
      pending[np++] = c;
      pending_units += cu;
    }
  }
  if (np) return launch_stream_chunks(pending, np, opt, tl_slot, st);
  return TFR_OK;
}

extern "C" int tfr_adam_slice_multi(const tfr_slice_update* sides, int32_t n_sides, int32_t width, int64_t n,
                                    const tfr_opt_scalars* opt, int32_t sgd, int32_t tl_slot, void* stream) {
  TFR_CHECK_ARG(sides && n_sides >= 1 && n_sides <= 2 && width > 0 && n >= 0 && tl_slot >= 0 && tl_slot < TFR_TL_SLOTS);
  if (n == 0) return TFR_OK;
  TFR_CHECK_ARG(sgd || opt);
  SliceSide ss[2];
  memset(ss, 0, sizeof(ss));
  for (int i = 0; i < n_sides; ++i) {
    const tfr_slice_update& u = sides[i];
    TFR_CHECK_ARG(u.sorted_ids);
    TFR_CHECK_ARG(!u.var || (u.gsum && (sgd || (u.m && u.v))));
    TFR_CHECK_ARG(!u.bvar || (u.bgsum && (sgd || (u.bm && u.bv))));
    ss[i].var = u.var; ss[i].m = u.m; ss[i].v = u.v; ss[i].bvar = u.bvar; ss[i].bm = u.bm; ss[i].bv = u.bv;
    ss[i].sid = u.sorted_ids; ss[i].gsum = u.gsum; ss[i].bgsum = u.bgsum;
  }
  const RowGeom g = row_geom(width);
  const int64_t groups = (n + SLICE_ENT - 1) / SLICE_ENT;
  const int groups_per_cta = 256 / g.lanes;
  dim3 grid((unsigned)((groups + groups_per_cta - 1) / groups_per_cta), (unsigned)n_sides);
  cudaStream_t st = (cudaStream_t)stream;
#define TFR_SLICE_CASE(V, LL)                                                                        \
  if (g.vec == V && g.lanes == LL) {                                                                 \
    TFR_PREP((adam_slice_kernel<V, LL>));                                                            \
    adam_slice_kernel<V, LL><<<grid, 256, 0, st>>>(ss[0], ss[1], width, n, opt, sgd, tl_slot);       \
    TFR_LAUNCH_CHECK();                                                                               \
    return TFR_OK;                                                                                    \
  }
  TFR_SLICE_CASE(4, 1) TFR_SLICE_CASE(4, 2) TFR_SLICE_CASE(4, 4) TFR_SLICE_CASE(4, 8) TFR_SLICE_CASE(4, 16)
  TFR_SLICE_CASE(4, 32) TFR_SLICE_CASE(1, 1) TFR_SLICE_CASE(1, 2) TFR_SLICE_CASE(1, 4) TFR_SLICE_CASE(1, 8)
  TFR_SLICE_CASE(1, 16) TFR_SLICE_CASE(1, 32)
#undef TFR_SLICE_CASE
  set_error("unsupported width %d", width);
  return TFR_ERR_INVALID;
}

extern "C" int tfr_adam_touched(float* var, float* m, float* v, int32_t width, const int32_t* sorted_ids, int64_t n,
                                const float* gsum, const tfr_opt_scalars* opt, void* stream) {
  tfr_slice_update u{var, m, v, nullptr, nullptr, nullptr, sorted_ids, gsum, nullptr};
  return tfr_adam_slice_multi(&u, 1, width, n, opt, 0, TFR_TL_SLOTS - 1, stream);
}

extern "C" int tfr_sgd_apply(float* var, int32_t width, const int32_t* sorted_ids, int64_t n, const float* gsum,
                             void* stream) {
  tfr_slice_update u{var, nullptr, nullptr, nullptr, nullptr, nullptr, sorted_ids, gsum, nullptr};
  return tfr_adam_slice_multi(&u, 1, width, n, nullptr, 1, TFR_TL_SLOTS - 1, stream);
}

extern "C" int tfr_svd_finish_step(const tfr_svd_tables* t, tfr_opt_scalars* opt, const int32_t* users,
                                   const int32_t* items, int64_t B, const tfr_svd_step_ws* ws, int32_t n_partials,
                                   void* stream) {
  TFR_CHECK_ARG(t && opt && users && items && ws && B > 0 && n_partials > 0 && n_partials <= TFR_MAX_PARTIALS);
  TFR_PREP(finish_step_kernel);
  finish_step_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*t, opt, users, items, B,
                                                                                  ws->partials, ws->se_partials,
                                                                                  n_partials);
  TFR_LAUNCH_CHECK();
  return TFR_OK;
}
