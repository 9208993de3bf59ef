// K3b: ordered segment sums of the per-occurrence gradients -- the "segment-reduce" half of the dedup
// (TF: tf.unsorted_segment_sum inside optimizer.py::_deduplicate_indexed_slices), fused
// with the backward itself so the [B,dim] per-occurrence gradient rows of TF's IndexedSlices are never
// written to HBM.
//
// Backward restated (TF autodiff of ops.py:124-126,140 through ops.py:44-47,81-89):
//   user row u, occurrence b:  g = e_b * v'_b + reg * u        (v' = |v| if TFR_ABS_ITEM)
//   item row v, occurrence b:  g = (e_b * u_b) [* sign(v)] + reg * v
//   bias:                      g = e_b (+ reg * bias if TFR_REG_BIAS)
// each product and the add a separate fp32 rounding (separate TF kernels + AddN).  In SGD mode
// (ops.py:145) the summand is lr*g so that the apply is a single subtraction.
//
// Work decomposition: the sorted (id,pos) array is cut into tiles of 32 entries, one lane group per
// tile.  A group walks its tile in order, keeping the running sum of the current run of equal ids.
// Runs that begin and end inside the tile are final: gsum[head index] = sum.  A run that crosses a
// tile boundary leaves a partial (cont[t]: continues a run begun earlier; tail[t]: begins here, goes
// on) and the fix-up kernel adds the partials of consecutive tiles in order.  The result is a fixed
// function of the input (deterministic) and equals the in-order sum except for regrouping at the
// 32-entry tile boundaries.
#include "common.cuh"

namespace tfr {

constexpr int SEG_TILE = 32;

struct SegSide {
  const int32_t* sid;      // sorted ids of this table
  const int32_t* spos;     // batch positions
  const int32_t* partner;  // the OTHER id column of the batch (items for the user table)
  const float* own_feat;   // this table's rows
  const float* partner_feat;
  const float* own_bias;
  float *gsum, *gsum_b, *cont, *cont_b, *tail, *tail_b;
  int is_item;
};

template <int VEC>
struct Acc {
  float v[VEC];
};

template <int VEC>
__device__ __forceinline__ Acc<VEC> load_units(const float* row, int unit) {
  Acc<VEC> a;
  if constexpr (VEC == 4) {
    const float4 x = ld_gather_f4(reinterpret_cast<const float4*>(row) + unit);
    a.v[0] = x.x; a.v[1] = x.y; a.v[2] = x.z; a.v[3] = x.w;
  } else {
    a.v[0] = ld_gather_f1(row + unit);
  }
  return a;
}

template <int VEC>
__device__ __forceinline__ void store_units(float* row, int unit, const Acc<VEC>& a) {
  if constexpr (VEC == 4) {
    reinterpret_cast<float4*>(row)[unit] = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
  } else {
    row[unit] = a.v[0];
  }
}

// UNITS = per-lane units (of VEC floats) needed to cover a row with L lanes.
template <int VEC, int L, int UNITS>
__global__ void __launch_bounds__(256) segsum_tiles_kernel(SegSide su, SegSide si, const tfr_opt_scalars* __restrict__ opt,
                                                           const float* __restrict__ err, int64_t B, int dim,
                                                           int n_tiles) {
  const SegSide s = blockIdx.y ? si : su;
  const int lane = threadIdx.x & (L - 1);
  const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  // whole lane groups leave together; L divides 32 so a warp may be partially active, all shuffles
  // below are confined to the group (width = L) and use the group's own mask.
  if (t >= n_tiles) return;
  const unsigned gmask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << ((threadIdx.x & 31) & ~(L - 1)));
  const int n_units = dim / VEC;
  const int flags = opt->flags;
  const float reg = opt->reg;
  const bool abs_item = flags & TFR_ABS_ITEM;
  const bool sgd = flags & TFR_OPT_SGD;
  const bool reg_bias = flags & TFR_REG_BIAS;
  const float lr = opt->lr;

  const int64_t k0 = t * SEG_TILE;
  const int64_t k1 = min(k0 + SEG_TILE, B);
  const int32_t prev_id = k0 > 0 ? s.sid[k0 - 1] : -1;
  const int32_t next_id = k1 < B ? s.sid[k1] : -1;

  Acc<VEC> acc[UNITS], own[UNITS];
  float acc_b = 0.0f, own_b = 0.0f;
  int32_t cur = -1;
  int64_t run_start = k0;

  auto flush = [&](int64_t k_end) {  // the run [run_start, k_end) of id `cur` is over (within this tile)
    const bool starts = run_start > k0 || cur != prev_id;
    const bool ends = k_end < k1 || cur != next_id;
    float* dst;
    float* dst_b;
    if (starts && ends) { dst = s.gsum + (size_t)run_start * dim; dst_b = s.gsum_b + run_start; }
    else if (!starts)   { dst = s.cont + (size_t)t * dim;         dst_b = s.cont_b + t; }
    else                { dst = s.tail + (size_t)t * dim;         dst_b = s.tail_b + t; }
#pragma unroll
    for (int q = 0; q < UNITS; ++q) {
      const int unit = lane + q * L;
      if (unit < n_units) store_units<VEC>(dst, unit, acc[q]);
    }
    if (lane == 0) *dst_b = acc_b;
  };

  for (int64_t kb = k0; kb < k1; kb += L) {
    // lane-parallel metadata fetch for up to L entries
    const int64_t k = kb + lane;
    int32_t my_id = -1, my_partner = 0;
    float my_e = 0.0f;
    if (k < k1) {
      my_id = s.sid[k];
      const int32_t b = s.spos[k];
      my_e = err[b];
      my_partner = s.partner[b];
    }
    const int cnt = (int)min((int64_t)L, k1 - kb);
    for (int j = 0; j < cnt; ++j) {
      const int32_t id = __shfl_sync(gmask, my_id, j, L);
      const float e = __shfl_sync(gmask, my_e, j, L);
      const int32_t pid = __shfl_sync(gmask, my_partner, j, L);
      if (id != cur) {
        if (cur >= 0) flush(kb + j);
        cur = id;
        run_start = kb + j;
        const float* orow = s.own_feat + (size_t)id * dim;
#pragma unroll
        for (int q = 0; q < UNITS; ++q) {
          const int unit = lane + q * L;
          if (unit < n_units) own[q] = load_units<VEC>(orow, unit);
#pragma unroll
          for (int c = 0; c < VEC; ++c) acc[q].v[c] = 0.0f;
        }
        own_b = reg_bias ? ld_gather_f1(s.own_bias + id) : 0.0f;
        acc_b = 0.0f;
      }
      const float* prow = s.partner_feat + (size_t)pid * dim;
#pragma unroll
      for (int q = 0; q < UNITS; ++q) {
        const int unit = lane + q * L;
        if (unit < n_units) {
          const Acc<VEC> p = load_units<VEC>(prow, unit);
#pragma unroll
          for (int c = 0; c < VEC; ++c) {
            float g;
            if (!s.is_item) {
              g = mul_rn(e, abs_item ? fabsf(p.v[c]) : p.v[c]);
            } else {
              g = mul_rn(e, p.v[c]);
              if (abs_item) {
                const float o = own[q].v[c];
                g = mul_rn(g, (o > 0.0f) ? 1.0f : ((o < 0.0f) ? -1.0f : 0.0f));
              }
            }
            g = add_rn(g, mul_rn(reg, own[q].v[c]));
            if (sgd) g = mul_rn(lr, g);
            acc[q].v[c] = add_rn(acc[q].v[c], g);
          }
        }
      }
      float gb = reg_bias ? add_rn(e, mul_rn(reg, own_b)) : e;
      if (sgd) gb = mul_rn(lr, gb);
      acc_b = add_rn(acc_b, gb);
    }
  }
  if (cur >= 0) flush(k1);
}

// One lane group per tile; only the tile in which a boundary-crossing run BEGINS does work: it adds
// tail[t] + cont[t+1] + cont[t+2] + ... in tile order and writes gsum at the run's head index.
template <int VEC, int L, int UNITS>
__global__ void __launch_bounds__(256) segsum_fixup_kernel(SegSide su, SegSide si, int64_t B, int dim, int n_tiles) {
  const SegSide s = blockIdx.y ? si : su;
  const int lane = threadIdx.x & (L - 1);
  const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  if (t >= n_tiles) return;
  const int n_units = dim / VEC;
  const int64_t k0 = t * SEG_TILE;
  const int64_t k1 = min(k0 + SEG_TILE, B);
  if (k1 >= B) return;  // last tile: nothing continues past it
  const int32_t last_id = s.sid[k1 - 1];
  if (s.sid[k1] != last_id) return;  // last run ends here
  // head of the last run inside this tile (ids are sorted: the run is a suffix of the tile)
  int64_t a = k1 - 1;
  while (a > k0 && s.sid[a - 1] == last_id) --a;
  if (a == k0 && k0 > 0 && s.sid[k0 - 1] == last_id) return;  // begun in an earlier tile
  Acc<VEC> tot[UNITS];
#pragma unroll
  for (int q = 0; q < UNITS; ++q) {
    const int unit = lane + q * L;
    if (unit < n_units) tot[q] = load_units<VEC>(s.tail + (size_t)t * dim, unit);
  }
  float tot_b = s.tail_b[t];
  for (int64_t tt = t + 1; tt < n_tiles; ++tt) {
#pragma unroll
    for (int q = 0; q < UNITS; ++q) {
      const int unit = lane + q * L;
      if (unit < n_units) {
        const Acc<VEC> c = load_units<VEC>(s.cont + (size_t)tt * dim, unit);
#pragma unroll
        for (int cc = 0; cc < VEC; ++cc) tot[q].v[cc] = add_rn(tot[q].v[cc], c.v[cc]);
      }
    }
    tot_b = add_rn(tot_b, s.cont_b[tt]);
    const int64_t e1 = min((tt + 1) * (int64_t)SEG_TILE, B);
    if (e1 >= B || s.sid[e1] != last_id) break;
  }
#pragma unroll
  for (int q = 0; q < UNITS; ++q) {
    const int unit = lane + q * L;
    if (unit < n_units) store_units<VEC>(s.gsum + (size_t)a * dim, unit, tot[q]);
  }
  if (lane == 0) s.gsum_b[a] = tot_b;
}

}  // namespace tfr

using namespace tfr;

extern "C" int tfr_svd_segment_grads(const tfr_svd_tables* t, const tfr_opt_scalars* opt, const int32_t* users,
                                     const int32_t* items, int64_t B, const tfr_svd_step_ws* ws, void* stream) {
  TFR_CHECK_ARG(t && opt && users && items && ws && B > 0 && t->dim > 0);
  const int dim = t->dim;
  const RowGeom g = row_geom(dim);
  const int units = (dim / g.vec + g.lanes - 1) / g.lanes;
  const int n_tiles = (int)((B + SEG_TILE - 1) / SEG_TILE);
  SegSide su{ws->su_ids, ws->su_pos, items, t->user_feat, t->item_feat, t->user_bias,
             ws->gsum_uf, ws->gsum_ub, ws->cont_uf, ws->cont_ub, ws->tail_uf, ws->tail_ub, 0};
  SegSide si{ws->si_ids, ws->si_pos, users, t->item_feat, t->user_feat, t->item_bias,
             ws->gsum_if, ws->gsum_ib, ws->cont_if, ws->cont_ib, ws->tail_if, ws->tail_ib, 1};
  const int groups_per_cta = 256 / g.lanes;
  dim3 grid((unsigned)((n_tiles + groups_per_cta - 1) / groups_per_cta), 2);
  cudaStream_t st = (cudaStream_t)stream;
#define TFR_SEG_CASE(V, LL, UU)                                                                                   \
  if (g.vec == V && g.lanes == LL && units == UU) {                                                               \
    segsum_tiles_kernel<V, LL, UU><<<grid, 256, 0, st>>>(su, si, opt, ws->err, B, dim, n_tiles);                  \
    TFR_LAUNCH_CHECK();                                                                                            \
    segsum_fixup_kernel<V, LL, UU><<<grid, 256, 0, st>>>(su, si, B, dim, n_tiles);                                \
    TFR_LAUNCH_CHECK();                                                                                            \
    return TFR_OK;                                                                                                 \
  }
  TFR_SEG_CASE(4, 1, 1) TFR_SEG_CASE(4, 2, 1) TFR_SEG_CASE(4, 4, 1) TFR_SEG_CASE(4, 8, 1) TFR_SEG_CASE(4, 16, 1)
  TFR_SEG_CASE(4, 32, 1) TFR_SEG_CASE(4, 32, 2) TFR_SEG_CASE(4, 32, 4)
  TFR_SEG_CASE(1, 1, 1) TFR_SEG_CASE(1, 2, 1) TFR_SEG_CASE(1, 4, 1) TFR_SEG_CASE(1, 8, 1) TFR_SEG_CASE(1, 16, 1)
  TFR_SEG_CASE(1, 32, 1) TFR_SEG_CASE(1, 32, 2) TFR_SEG_CASE(1, 32, 4)
#undef TFR_SEG_CASE
  set_error("unsupported dim %d (vec %d lanes %d units %d)", dim, g.vec, g.lanes, units);
  return TFR_ERR_INVALID;
}
