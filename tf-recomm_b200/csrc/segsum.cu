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
  uint8_t* kind;  // per tile: TILE_MID | TILE_START (see segsum_fixup_kernel)
  int32_t* slot;  // [rows] row -> head index of its run (where its gsum lives); -1 between steps
  int32_t n_rows; // ids >= n_rows mark occurrences owned by another rank (row-sharded mode): skipped
  // factorization-machine mode (xval != null): the sorted ids are FEATURE ids of the batch's non-zeros, position =
  // index p of the non-zero; partner row = sums[rowof[p]] (the CSR row's sum_i V_i x_i), and
  //   g_V = e_r * (x_p * (sums_r - V_f * x_p)) + reg * V_f        (Rendle 2010 eq. 4; forward.py:21-22's model)
  //   g_W = e_r * x_p (+ reg * W_f if TFR_REG_BIAS)
  const float* xval;      // [nnz] feature values
  const int32_t* rowof;   // [nnz] CSR row of every non-zero
  int is_item;
};

// tile classification written by the tiles kernel, read by the fix-up:
//   TILE_MID   the whole tile is ONE run that began in an earlier tile and goes on into the next
//   TILE_START the tile's last run begins here and goes on into the next tile (tail[t] is valid)
constexpr uint8_t TILE_MID = 1, TILE_START = 2;

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
  TlScope tl_scope(opt, TFR_TL_TILES);
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
  if (s.sid[k0] >= s.n_rows) {  // the whole tile belongs to other ranks
    if (lane == 0) s.kind[t] = 0;
    return;
  }

  Acc<VEC> acc[UNITS], own[UNITS];
  float acc_b = 0.0f, own_b = 0.0f;
  int32_t cur = -1;
  int64_t run_start = k0;
  uint8_t kind = 0;

  auto flush = [&](int64_t k_end) {  // the run [run_start, k_end) of id `cur` is over (within this tile)
    const bool starts = run_start > k0 || cur != prev_id;
    const bool ends = k_end < k1 || cur != next_id;
    float* dst;
    float* dst_b;
    if (starts && ends) { dst = s.gsum + (size_t)run_start * dim; dst_b = s.gsum_b + run_start; }
    else if (!starts)   { dst = s.cont + (size_t)t * dim;         dst_b = s.cont_b + t; if (!ends) kind |= TILE_MID; }
    else                { dst = s.tail + (size_t)t * dim;         dst_b = s.tail_b + t; kind |= TILE_START; }
#pragma unroll
    for (int q = 0; q < UNITS; ++q) {
      const int unit = lane + q * L;
      if (unit < n_units) store_units<VEC>(dst, unit, acc[q]);
    }
    if (lane == 0) {
      *dst_b = acc_b;
      if (starts && ends) s.slot[cur] = (int32_t)run_start;
    }
  };

  for (int64_t kb = k0; kb < k1; kb += L) {
    // lane-parallel metadata fetch for up to L entries
    const int64_t k = kb + lane;
    int32_t my_id = -1, my_partner = 0;
    float my_e = 0.0f, my_x = 1.0f;
    if (k < k1) {
      my_id = s.sid[k];
      const int32_t b = s.spos[k];
      if (s.xval) {  // FM: b is a non-zero; its error and "partner" (the row's sums) come from its CSR row
        my_partner = s.rowof[b];
        my_e = err[my_partner];
        my_x = s.xval[b];
      } else {
        my_e = err[b];
        my_partner = s.partner ? s.partner[b] : b;  // row-sharded mode: partner rows are gathered by position
      }
    }
    const int cnt = (int)min((int64_t)L, k1 - kb);
    for (int j = 0; j < cnt; ++j) {
      const int32_t id = __shfl_sync(gmask, my_id, j, L);
      const float e = __shfl_sync(gmask, my_e, j, L);
      const int32_t pid = __shfl_sync(gmask, my_partner, j, L);
      const float xv = __shfl_sync(gmask, my_x, j, L);
      if (id >= s.n_rows) {  // sorted to the end: nothing of mine follows
        if (cur >= 0) flush(kb + j);
        cur = -1;
        kb = k1;  // leave both loops
        break;
      }
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
            if (s.xval) {
              const float tt = sub_rn(p.v[c], mul_rn(own[q].v[c], xv));
              g = mul_rn(e, mul_rn(xv, tt));
            } else if (!s.is_item) {
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
      float gb = s.xval ? mul_rn(e, xv) : e;
      if (reg_bias) gb = add_rn(gb, mul_rn(reg, own_b));
      if (sgd) gb = mul_rn(lr, gb);
      acc_b = add_rn(acc_b, gb);
    }
  }
  if (cur >= 0) flush(k1);
  if (lane == 0) s.kind[t] = kind;
}

// Fix-up: one CTA per tile, only CTAs of TILE_START tiles work.  A run that begins in tile t0 and ends in
// tile t1 has the partial sums tail[t0], cont[t0+1], ..., cont[t1]; tiles t0+1..t1-1 are TILE_MID.  The CTA's G
// thread groups add the cont rows j = g, g+G, g+2G, ... (each in increasing j), then group 0 adds
// tail + p_0 + p_1 + ... + p_{G-1}: a fixed tree, so the result is deterministic, and a hot row with
// thousands of occurrences costs ~n/(32*G) dependent steps instead of n/32.
template <int VEC>
__global__ void __launch_bounds__(256) segsum_fixup_kernel(SegSide su, SegSide si, const tfr_opt_scalars* __restrict__ opt,
                                                           int64_t B, int dim, int n_tiles, int cw) {
  TlScope tl_scope(opt, TFR_TL_FIXUP);
  extern __shared__ float s_part[];  // [G][dim] (+ [G] bias partials)
  const SegSide s = blockIdx.y ? si : su;
  const int t0 = blockIdx.x;
  if (!(s.kind[t0] & TILE_START)) return;
  const int G = 256 / cw;
  const int g = threadIdx.x / cw, c = threadIdx.x % cw;
  const int n_units = dim / VEC;
  __shared__ int s_t1;
  if (threadIdx.x < 32) {  // t1 = first tile after t0 that is not TILE_MID
    int t1 = -1;
    for (int base = t0 + 1; t1 < 0; base += 32) {
      const int tt = base + threadIdx.x;
      const bool stop = tt >= n_tiles || !(s.kind[tt] & TILE_MID);
      const unsigned m = __ballot_sync(0xffffffffu, stop);
      if (m) t1 = base + __ffs(m) - 1;
    }
    if (threadIdx.x == 0) s_t1 = min(t1, n_tiles - 1);
  }
  __syncthreads();
  const int t1 = s_t1;
  float* bias_part = s_part + (size_t)G * dim;
  for (int unit = c; unit < n_units; unit += cw) {
    Acc<VEC> acc;
#pragma unroll
    for (int q = 0; q < VEC; ++q) acc.v[q] = 0.0f;
    for (int tt = t0 + 1 + g; tt <= t1; tt += G) {
      const Acc<VEC> x = load_units<VEC>(s.cont + (size_t)tt * dim, unit);
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc.v[q] = add_rn(acc.v[q], x.v[q]);
    }
#pragma unroll
    for (int q = 0; q < VEC; ++q) s_part[(size_t)g * dim + unit * VEC + q] = acc.v[q];
  }
  if (c == 0) {
    float ab = 0.0f;
    for (int tt = t0 + 1 + g; tt <= t1; tt += G) ab = add_rn(ab, s.cont_b[tt]);
    bias_part[g] = ab;
  }
  __syncthreads();
  // head index of the run inside t0 (ids are sorted: the run is the tile's suffix)
  const int64_t k0 = (int64_t)t0 * SEG_TILE, k1 = min(k0 + SEG_TILE, B);
  const int32_t id = s.sid[k1 - 1];
  int64_t a = k1 - 1;
  while (a > k0 && s.sid[a - 1] == id) --a;
  for (int col = threadIdx.x; col < dim; col += 256) {
    float tot = s.tail[(size_t)t0 * dim + col];
    for (int gg = 0; gg < G; ++gg) tot = add_rn(tot, s_part[(size_t)gg * dim + col]);
    s.gsum[(size_t)a * dim + col] = tot;
  }
  if (threadIdx.x == 0) {
    float tot = s.tail_b[t0];
    for (int gg = 0; gg < G; ++gg) tot = add_rn(tot, bias_part[gg]);
    s.gsum_b[a] = tot;
    s.slot[id] = (int32_t)a;
  }
}

}  // namespace tfr

using namespace tfr;

static int launch_segsum(const SegSide& su, const SegSide& si, int n_sides, const tfr_opt_scalars* opt,
                         const float* err, int64_t B, int dim, cudaStream_t st) {
  const RowGeom g = row_geom(dim);
  const int units = (dim / g.vec + g.lanes - 1) / g.lanes;
  const int n_tiles = (int)((B + SEG_TILE - 1) / SEG_TILE);
  int cw = 1;
  while (cw < dim / g.vec && cw < 256) cw <<= 1;
  const int G = 256 / cw;
  const size_t fix_smem = ((size_t)G * dim + G) * sizeof(float);
  dim3 fix_grid((unsigned)n_tiles, (unsigned)n_sides);
  const int groups_per_cta = 256 / g.lanes;
  dim3 grid((unsigned)((n_tiles + groups_per_cta - 1) / groups_per_cta), (unsigned)n_sides);
#define TFR_SEG_CASE(V, LL, UU)                                                                                   \
  if (g.vec == V && g.lanes == LL && units == UU) {                                                               \
    TFR_PREP((segsum_tiles_kernel<V, LL, UU>));                                                                    \
    TFR_PREP((segsum_fixup_kernel<V>));                                                                            \
    segsum_tiles_kernel<V, LL, UU><<<grid, 256, 0, st>>>(su, si, opt, err, B, dim, n_tiles);                      \
    TFR_LAUNCH_CHECK();                                                                                            \
    segsum_fixup_kernel<V><<<fix_grid, 256, fix_smem, st>>>(su, si, opt, B, dim, n_tiles, cw);                    \
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

extern "C" int tfr_svd_segment_grads(const tfr_svd_tables* t, const tfr_opt_scalars* opt, const int32_t* users,
                                     const int32_t* items, int64_t B, const tfr_svd_step_ws* ws, void* stream) {
  TFR_CHECK_ARG(t && opt && users && items && ws && B > 0 && t->dim > 0 && t->user_slot && t->item_slot);
  const bool gathered = t->g_user_feat != nullptr;
  TFR_CHECK_ARG(!gathered || t->g_item_feat);
  SegSide su{ws->su_ids, ws->su_pos, gathered ? nullptr : items, t->user_feat,
             gathered ? t->g_item_feat : t->item_feat, t->user_bias,
             ws->gsum_uf, ws->gsum_ub, ws->cont_uf, ws->cont_ub, ws->tail_uf, ws->tail_ub, ws->kind_u, t->user_slot,
             t->user_num, nullptr, nullptr, 0};
  SegSide si{ws->si_ids, ws->si_pos, gathered ? nullptr : users, t->item_feat,
             gathered ? t->g_user_feat : t->user_feat, t->item_bias,
             ws->gsum_if, ws->gsum_ib, ws->cont_if, ws->cont_ib, ws->tail_if, ws->tail_ib, ws->kind_i, t->item_slot,
             t->item_num, nullptr, nullptr, 1};
  return launch_segsum(su, si, 2, opt, ws->err, B, t->dim, (cudaStream_t)stream);
}

// FM: one table of feature rows V [n_feat, dim] (+ linear weights W [n_feat]); the "batch" of the segment sums is
// the batch's nnz non-zeros.  ws must be carved for B = nnz (the user-side buffers are used).
extern "C" int tfr_fm_segment_grads(const float* V, const float* W, int32_t* slot, int32_t n_feat, int32_t dim,
                                    const tfr_opt_scalars* opt, const float* sums, const float* err,
                                    const float* xval, const int32_t* rowof, int64_t nnz, const tfr_svd_step_ws* ws,
                                    void* stream) {
  TFR_CHECK_ARG(V && W && slot && opt && sums && err && xval && rowof && ws && nnz > 0 && dim > 0 && n_feat > 0);
  SegSide sf{ws->su_ids, ws->su_pos, nullptr, V, sums, W, ws->gsum_uf, ws->gsum_ub, ws->cont_uf, ws->cont_ub,
             ws->tail_uf, ws->tail_ub, ws->kind_u, slot, n_feat, xval, rowof, 0};
  return launch_segsum(sf, sf, 1, opt, err, nnz, dim, (cudaStream_t)stream);
}
