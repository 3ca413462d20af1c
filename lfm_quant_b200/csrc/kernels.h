// Host-callable launchers of the lfmq kernels.  All return 0 or LFMQ_ERR_CUDA (message in g_err).
#pragma once
#include "common.cuh"

namespace lfmq {

// ---- fp32 SIMT building blocks (kernels_simt.cu) ---------------------------------------------
// C[M,N] (row-major, ldc) = A*B + beta*C with A(m,k) = A[m*sAm + k*sAk], B(k,n) = B[k*sBk + n*sBn].
// `scratch` (>= scratch_elems floats) is used for deterministic split-K partials.
int sgemm(cudaStream_t s, int M, int N, int K, const float* A, long sAm, long sAk, const float* B, long sBk,
          long sBn, float* C, long ldc, float beta, float* scratch, size_t scratch_elems);

int lstm_pointwise_fwd(cudaStream_t s, int B, int T, int H, int t, const float* z, const float* bias, float* gates,
                       float* c, float* h, const float* rmask, float* hm);
int lstm_pointwise_bwd(cudaStream_t s, int B, int T, int H, int t, const float* gates, const float* c,
                       const float* dh_out, const float* dh_rec, const float* rmask, float* dc, float* dz);
int gru_pointwise_fwd(cudaStream_t s, int B, int T, int H, int t, const float* zx, const float* zh, const float* bias,
                      float* gates, float* h, const float* rmask, float* hm);
int gru_pointwise_bwd(cudaStream_t s, int B, int T, int H, int t, const float* gates, const float* h,
                      const float* dh_out, const float* dh_rec, const float* rmask, float* dcarry, float* dxz,
                      float* dhz);
int gen_row_mask(cudaStream_t s, int B, int H, DropoutKey key, int64_t row0, float* rmask);
int shift_mask(cudaStream_t s, int B, int T, int H, const float* h, const float* rmask, float* hp);
int bn_dropout_fwd(cudaStream_t s, int B, int T, int H, const float* h, const float* gamma, const float* beta,
                   const float* mean, const float* var, float eps, bool use_dropout, DropoutKey key, int64_t row0,
                   float* y);
int bn_dropout_bwd(cudaStream_t s, int B, int T, int H, const float* dy, const float* h, const float* gamma,
                   const float* mean, const float* var, float eps, bool use_dropout, DropoutKey key, int64_t row0,
                   float* dh_out, float* dgamma, float* dbeta, float* scratch, size_t scratch_elems);
int add_bias_rows(cudaStream_t s, long rows, int N, float* C, const float* bias);
int colsum(cudaStream_t s, long rows, int N, const float* A, float* out, float* scratch, size_t scratch_elems);
int softplus_floor(cudaStream_t s, long n, const float* a, float* var);
int uq_loss_grad(cudaStream_t s, int B, int T, int O, const float* pred, const float* var, const float* apre,
                 const float* y, const float* counts_in, int target_idx, float p1, float p2, float* dpred, float* da,
                 float* out_loss, float* out_uq0, float* out_mse0, float* counts_out, float* scratch);
int mask_count(cudaStream_t s, int B, int T, int O, const float* y, float* out2, unsigned int* tickets);
// out2 = {loss, mse_0}; maskout2 (nullable) = {B, mask_count_local}; dpred may be NULL (validation);
// denom (nullable) = device {B_global, mask_count_global}.
int loss_grad(cudaStream_t s, int B, int T, int O, const float* pred, const float* y, const float* denom,
              int target_idx, float p1, float p2, float* dpred, float* out2, float* maskout2, float* scratch);
// scalars[0] = ||g||, scalars[1] = clip scale (1 when clip <= 0)
int grad_norm_scale(cudaStream_t s, long n, const float* g, float clip, float* scalars, float* scratch,
                    unsigned int* ticket);
int opt_update(cudaStream_t s, int opt, long n, float* p, const float* g, float* slot0, float* slot1,
               const float* scalars, float lr, float c1, float c2, float momentum);
int maxnorm_cols(cudaStream_t s, int I, int N, float* W, float max_norm);
int fill(cudaStream_t s, float* p, long n, float v);

// forecast_steps > 1 (rnn_point_estimate.py:109-150): glue between the stage handles
constexpr int LFMQ_MAX_STAGES = 8;
struct ChainPtrs { float* p[LFMQ_MAX_STAGES]; };
struct ChainWeights { float w[LFMQ_MAX_STAGES]; };
int chain_next_input(cudaStream_t s, int B, int T, int F, int O, const float* prev, const float* pred, const float* x0,
                     float* next);
int scale_inplace(cudaStream_t s, long n, float* p, float w);
int chain_scatter_dx(cudaStream_t s, int B, int T, int F, int O, int stage, const float* dx, const ChainPtrs& dpred);
int chain_combine(cudaStream_t s, int S, const ChainPtrs& scalars, float clip, const ChainPtrs& loss2,
                  const ChainWeights& w, float* out2);

struct GatherArgs {
  int n_rows, n_cols, B, T, F, O, stride, seq_norm_col, log_squasher, aux_masking;
  const double* table;
  const int32_t* inp_idx;
  const int32_t* tar_idx;
  const int32_t* inp_cols;
  const int32_t* fin_cols;
  const double* center;
  const double* scale;
  const uint8_t* scale_flag;
  const uint8_t* aux_flag;
  float* x;
  float* y;
  double* seq_norm;
};
int gather_batch(cudaStream_t s, const GatherArgs& a);

// window index on the device (data_processing.py:170-305); work: 4 * ceil(n / 1024) + n ints
struct WindowIndexArgs {
  int n, train, stride, forecast_n, min_steps, max_steps;
  int32_t start_date, end_date, last_train_date;
  const int32_t* key;
  const uint8_t* active;
  const int32_t* date;
};
int window_index(cudaStream_t s, const WindowIndexArgs& a, int cap, int32_t* inp, int32_t* tar, int32_t* rows,
                 int32_t* count, int* work);
int unscale(cudaStream_t s, const float* in, float* out, long n_rows, int O, const double* scale, const double* center,
            int log_squasher);

}  // namespace lfmq
