// K0 batch assembly, K1 forward (eval), K2 forward + d cost/d logits (train).
//
// Reference arithmetic replaced here (file:line into /root/reference):
//   dataio.py:114-117   ShuffleIterator.next: rows = inputs[randint ids]           -> batch_assemble_kernel
//   ops.py:13-14,37-38  four embedding_lookup gathers                               -> row loads below
//   ops.py:44-47        logits = ((sum_k u*v' + mu) + b_u) + b_i                     -> svd_logit
//   ops.py:76-78        fork head: infer = round(sigmoid(logits)); README: logits   -> head()
//   ops.py:124-126      d cost/d logits (squared error / sigmoid-CE)                 -> tfr::dloss
//
// Layout: warp lanes are split in groups of L lanes, one group per batch row; a lane owns VEC
// consecutive floats of the row per pass (128-bit loads when dim % 4 == 0).  Four rows per group are
// in flight (8 gathers outstanding per lane).  Products and adds are separate fp32 roundings
// (tf.multiply then tf.reduce_sum), the lane-group butterfly fixes the reduction order.
#include <string.h>

#include "common.cuh"

namespace tfr {

__device__ __forceinline__ float head(int flags, float x) {
  return (flags & TFR_LOSS_SIGMOID_CE) ? rintf(sigmoid_tf(x)) : x;
}

// TRAIN = false: eval forward.  TRAIN = true: also err[b], per-CTA partial sums of err and of the
// float64 squared error (svd_train_val.py:104).
template <int VEC, int L, bool TRAIN>
__global__ void __launch_bounds__(256) svd_forward_kernel(tfr_svd_tables t, const tfr_opt_scalars* __restrict__ opt,
                                                          const int32_t* __restrict__ users,
                                                          const int32_t* __restrict__ items,
                                                          const float* __restrict__ rates, int64_t B, int flags_arg,
                                                          float* __restrict__ logits, float* __restrict__ infer,
                                                          float* __restrict__ err, float* __restrict__ partials,
                                                          double* __restrict__ se_partials) {
  TlScope tl_scope(TRAIN ? opt : nullptr, TFR_TL_FWD);
  constexpr int GPW = 32 / L;  // groups per warp
  const int lane = threadIdx.x & (L - 1);
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  const int64_t n_groups = (int64_t)gridDim.x * blockDim.x / L;
  const int flags = TRAIN ? opt->flags : flags_arg;
  const bool abs_item = flags & TFR_ABS_ITEM;
  const int dim = t.dim;
  const float mu = *t.mu;
  const bool gathered = t.g_user_feat != nullptr;
  const float* __restrict__ uf = gathered ? t.g_user_feat : t.user_feat;
  const float* __restrict__ itf = gathered ? t.g_item_feat : t.item_feat;
  const float* __restrict__ ubias = gathered ? t.g_user_bias : t.user_bias;
  const float* __restrict__ ibias = gathered ? t.g_item_bias : t.item_bias;
  const size_t fs = gathered ? (t.g_stride ? (size_t)t.g_stride : (size_t)dim)
                             : (t.feat_stride == 0 ? (size_t)dim : (size_t)t.feat_stride);  // floats between rows
  float err_acc = 0.0f;
  double se_acc = 0.0;

  // all lanes of a warp iterate the same number of times (the butterflies need full warps).  R rows per group
  // are in flight: their 2*R row loads (and lane 0's bias / rating loads) are issued before any arithmetic.
  constexpr int R = 4;
  const int64_t warp_first = group - (group % GPW);
  for (int64_t b0 = warp_first; b0 < B; b0 += R * n_groups) {
    int64_t bs[R];
    bool val[R];
    int32_t uu[R], ii[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      bs[r] = b0 + (group % GPW) + r * n_groups;
      val[r] = bs[r] < B;
      // row-sharded mode: the batch's rows were gathered by position (t.g_*), otherwise gather by id
      uu[r] = val[r] ? (gathered ? (int32_t)bs[r] : users[bs[r]]) : 0;
      ii[r] = val[r] ? (gathered ? (int32_t)bs[r] : items[bs[r]]) : 0;
    }
    float bu[R], bi[R], zz[R];
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        bu[r] = ld_gather_f1(ubias + uu[r]);
        bi[r] = ld_gather_f1(ibias + ii[r]);
        zz[r] = (TRAIN && val[r]) ? rates[bs[r]] : 0.0f;
      }
    }
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0f;
    if (VEC == 4) {
      const int n4 = dim >> 2;
      for (int k = lane; k < n4; k += L) {
        float4 a[R], q[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          a[r] = ld_gather_f4(reinterpret_cast<const float4*>(uf + (size_t)uu[r] * fs) + k);
          q[r] = ld_gather_f4(reinterpret_cast<const float4*>(itf + (size_t)ii[r] * fs) + k);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (abs_item) { q[r].x = fabsf(q[r].x); q[r].y = fabsf(q[r].y); q[r].z = fabsf(q[r].z); q[r].w = fabsf(q[r].w); }
          acc[r] = add_rn(acc[r], mul_rn(a[r].x, q[r].x));
          acc[r] = add_rn(acc[r], mul_rn(a[r].y, q[r].y));
          acc[r] = add_rn(acc[r], mul_rn(a[r].z, q[r].z));
          acc[r] = add_rn(acc[r], mul_rn(a[r].w, q[r].w));
        }
      }
    } else {
      for (int k = lane; k < dim; k += L) {
        float a[R], q[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          a[r] = ld_gather_f1(uf + (size_t)uu[r] * fs + k);
          q[r] = ld_gather_f1(itf + (size_t)ii[r] * fs + k);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = add_rn(acc[r], mul_rn(a[r], abs_item ? fabsf(q[r]) : q[r]));
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = group_sum<L>(acc[r]);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (!val[r]) continue;
        const int64_t b = bs[r];
        float x = add_rn(acc[r], mu);   // ops.py:45
        x = add_rn(x, bu[r]);           // ops.py:46
        x = add_rn(x, bi[r]);           // ops.py:47
        const float inf = head(flags, x);
        if (logits) logits[b] = x;
        if (infer) infer[b] = inf;
        if (TRAIN) {
          const float e = dloss(flags, x, zz[r]);
          err[b] = e;
          err_acc = add_rn(err_acc, e);
          const double dse = (double)zz[r] - (double)inf;
          se_acc += dse * dse;
        }
      }
    }
  }
  if (TRAIN) {
    // fixed-order block reduction: group leaders -> shared -> thread 0 sums sequentially.
    __shared__ float s_err[256];
    __shared__ double s_se[256];
    s_err[threadIdx.x] = (lane == 0) ? err_acc : 0.0f;
    s_se[threadIdx.x] = (lane == 0) ? se_acc : 0.0;
    __syncthreads();
    if (threadIdx.x < 32) {
      float a = 0.0f;
      double d = 0.0;
      for (int j = threadIdx.x; j < 256; j += 32) { a = add_rn(a, s_err[j]); d += s_se[j]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a = add_rn(a, __shfl_xor_sync(0xffffffffu, a, o));
        d += __shfl_xor_sync(0xffffffffu, d, o);
      }
      if (threadIdx.x == 0) { partials[blockIdx.x] = a; se_partials[blockIdx.x] = d; }
    }
  }
}

// lr_t = lr*sqrt(1-beta2_power)/(1-beta1_power), three fp32 roundings (TF: adam.py _apply_sparse_shared)
__device__ __forceinline__ void begin_step_scalars(tfr_opt_scalars* opt) {
  float tt = sqrt_rn(sub_rn(1.0f, opt->beta2_power));
  tt = mul_rn(opt->lr, tt);
  opt->lr_t = div_rn(tt, sub_rn(1.0f, opt->beta1_power));
}

__global__ void begin_step_kernel(tfr_opt_scalars* opt) { begin_step_scalars(opt); }

__global__ void __launch_bounds__(256) batch_assemble_kernel(tfr_svd_tables t, tfr_opt_scalars* opt,
                                                             const int32_t* __restrict__ col_user,
                                                             const int32_t* __restrict__ col_item,
                                                             const float* __restrict__ col_rate,
                                                             const int64_t* __restrict__ row_index, int64_t batch_index,
                                                             int64_t B, int32_t* __restrict__ users,
                                                             int32_t* __restrict__ items, float* __restrict__ rates) {
  TlScope tl_scope(opt, TFR_TL_ASSEMBLE);
  // batch_index >= 0: that batch; -1: the batch at batch_cursor (this step's, on the step's stream); -2: the batch at
  // prefetch_cursor (assembled ahead on a side stream: a counter the concurrent step never writes)
  const int64_t batch = batch_index >= 0 ? batch_index : (batch_index == -1 ? opt->batch_cursor : opt->prefetch_cursor);
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    const int64_t row = row_index[batch * B + b];
    users[b] = col_user[row];
    items[b] = col_item[row];
    rates[b] = col_rate[row];
  }
}

__global__ void advance_prefetch_cursor_kernel(tfr_opt_scalars* opt) { opt->prefetch_cursor += 1; }
__global__ void set_cursor_kernel(tfr_opt_scalars* opt, int64_t k) { opt->batch_cursor = k; opt->prefetch_cursor = k; }

__global__ void prime_prefetch_cursor_kernel(tfr_opt_scalars* opt) { opt->prefetch_cursor = opt->batch_cursor + 1; }

int advance_prefetch_cursor(tfr_opt_scalars* opt, cudaStream_t st, bool prime) {
  if (prime) prime_prefetch_cursor_kernel<<<1, 1, 0, st>>>(opt);
  else advance_prefetch_cursor_kernel<<<1, 1, 0, st>>>(opt);
  TFR_LAUNCH_CHECK();
  return TFR_OK;
}

template <bool TRAIN>
static int launch_forward(const tfr_svd_tables* t, const tfr_opt_scalars* opt, const int32_t* users,
                          const int32_t* items, const float* rates, int64_t B, int flags, float* logits, float* infer,
                          float* err, float* partials, double* se_partials, int* n_partials_out, cudaStream_t st) {
  const RowGeom g = row_geom(t->dim);
  const int64_t rows_per_cta = 256 / g.lanes * 4;
  int64_t grid = (B + rows_per_cta - 1) / rows_per_cta;
  const int64_t cap = TRAIN ? TFR_MAX_PARTIALS : (int64_t)sm_count() * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (n_partials_out) *n_partials_out = (int)grid;
#define TFR_FWD_CASE(V, LL)                                                                                     \
  if (g.vec == V && g.lanes == LL) {                                                                            \
    TFR_PREP((svd_forward_kernel<V, LL, TRAIN>));                                                               \
    svd_forward_kernel<V, LL, TRAIN><<<(unsigned)grid, 256, 0, st>>>(*t, opt, users, items, rates, B, flags, logits, \
                                                                     infer, err, partials, se_partials);        \
    TFR_LAUNCH_CHECK();                                                                                          \
    return TFR_OK;                                                                                               \
  }
  TFR_FWD_CASE(4, 1) TFR_FWD_CASE(4, 2) TFR_FWD_CASE(4, 4) TFR_FWD_CASE(4, 8) TFR_FWD_CASE(4, 16) TFR_FWD_CASE(4, 32)
  TFR_FWD_CASE(1, 1) TFR_FWD_CASE(1, 2) TFR_FWD_CASE(1, 4) TFR_FWD_CASE(1, 8) TFR_FWD_CASE(1, 16) TFR_FWD_CASE(1, 32)
#undef TFR_FWD_CASE
  set_error("unsupported dim %d", t->dim);
  return TFR_ERR_INVALID;
}

int fwd_err_n_partials(int dim, int64_t B) {
  const RowGeom g = row_geom(dim);
  const int64_t rows_per_cta = 256 / g.lanes * 4;
  int64_t grid = (B + rows_per_cta - 1) / rows_per_cta;
  if (grid > TFR_MAX_PARTIALS) grid = TFR_MAX_PARTIALS;
  if (grid < 1) grid = 1;
  return (int)grid;
}

}  // namespace tfr

using namespace tfr;

extern "C" int tfr_svd_forward(const tfr_svd_tables* t, const int32_t* users, const int32_t* items, int64_t B,
                               int32_t flags, float* logits, float* infer, void* stream) {
  TFR_CHECK_ARG(t && t->dim > 0 && B >= 0);
  if (B == 0) return TFR_OK;
  TFR_CHECK_ARG(users && items && t->user_feat && t->item_feat && t->user_bias && t->item_bias && t->mu);
  return launch_forward<false>(t, nullptr, users, items, nullptr, B, flags, logits, infer, nullptr, nullptr, nullptr,
                               nullptr, (cudaStream_t)stream);
}

extern "C" int tfr_svd_fwd_err(const tfr_svd_tables* t, const tfr_opt_scalars* opt, const int32_t* users,
                               const int32_t* items, const float* rates, int64_t B, float* logits, float* infer,
                               const tfr_svd_step_ws* ws, void* stream) {
  TFR_CHECK_ARG(t && opt && ws && t->dim > 0 && B > 0 && users && items && rates);
  return launch_forward<true>(t, opt, users, items, rates, B, 0, logits, infer, ws->err, ws->partials,
                              ws->se_partials, nullptr, (cudaStream_t)stream);
}

extern "C" int tfr_svd_batch_assemble(const tfr_svd_tables* t, tfr_opt_scalars* opt, const int32_t* col_user,
                                      const int32_t* col_item, const float* col_rate, const int64_t* row_index,
                                      int64_t batch_index, int64_t B, int32_t* users, int32_t* items, float* rates,
                                      void* stream) {
  TFR_CHECK_ARG(t && opt && B > 0 && users && items && rates && col_user && col_item && col_rate && row_index);
  TFR_CHECK_ARG(batch_index >= -2);
  TFR_PREP(batch_assemble_kernel);
  batch_assemble_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      *t, opt, col_user, col_item, col_rate, row_index, batch_index, B, users, items, rates);
  TFR_LAUNCH_CHECK();
  return TFR_OK;
}

extern "C" int tfr_opt_set_cursor(tfr_opt_scalars* opt_dev, int64_t k, void* stream) {
  TFR_CHECK_ARG(opt_dev && k >= 0);
  set_cursor_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(opt_dev, k);
  TFR_LAUNCH_CHECK();
  return TFR_OK;
}

extern "C" int tfr_svd_begin_step(tfr_opt_scalars* opt, void* stream) {
  TFR_CHECK_ARG(opt);
  begin_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(opt);
  TFR_LAUNCH_CHECK();
  return TFR_OK;
}
