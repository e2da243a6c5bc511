/*
 * gac_b200.h -- C ABI of libgac_b200.so: the B200 (sm_100a) implementation of the
 * Generative Actor Critic train step of gwbcho/dpo-replication.
 *
 * The reference has no FFI layer: its boundary is the Python class surface of
 * GAC/networks.py and GAC/agent.py.  Each entry point below names the reference
 * function (file:line under the reference repo) whose arithmetic it replaces; the
 * Python classes in dpo_replication_b200/ bind these with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory unless it says "host";
 *   - all matrices are fp32, row-major, densely packed unless a leading dimension is given;
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), performs
 *     no allocation and no host synchronisation;
 *   - returns 0 on success, a negative gac_status otherwise; gac_last_error() gives the text;
 *   - one context per device, not thread-safe per context (the reference is single-threaded).
 */
#ifndef GAC_B200_H
#define GAC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAC_B200_ABI_VERSION 4

enum gac_status {
  GAC_OK = 0,
  GAC_ERR_INVALID = -1,    /* bad argument */
  GAC_ERR_CUDA = -2,       /* a CUDA runtime call or launch failed */
  GAC_ERR_WORKSPACE = -3,  /* caller's workspace is too small */
  GAC_ERR_UNSUPPORTED = -4 /* e.g. mode not in {linear, boltzmann}: networks.py:94-95 */
};

/* variable order inside one flat parameter arena: actor (13), fnn1 (6), fnn2 (6), value (6).
 * The actor order is tf.Module's (attribute names sorted): action_embedding, dense_layer_1,
 * dense_layer_2, noise_embedding, rnn, state_embedding (networks.py:159-174). */
enum gac_var {
  GAC_A_WA = 0, GAC_A_BA, GAC_A_W1, GAC_A_B1, GAC_A_W2, GAC_A_B2, GAC_A_WN, GAC_A_BN,
  GAC_A_KX, GAC_A_U, GAC_A_GB, GAC_A_WS, GAC_A_BS,
  GAC_F1_W0, GAC_F1_B0, GAC_F1_W1, GAC_F1_B1, GAC_F1_W2, GAC_F1_B2,
  GAC_F2_W0, GAC_F2_B0, GAC_F2_W1, GAC_F2_B1, GAC_F2_W2, GAC_F2_B2,
  GAC_V_W0, GAC_V_B0, GAC_V_W1, GAC_V_B1, GAC_V_W2, GAC_V_B2,
  GAC_NVARS
};
enum gac_net { GAC_NET_ACTOR = 0, GAC_NET_FNN1, GAC_NET_FNN2, GAC_NET_VALUE, GAC_NNETS };

/* actor_kind: the AIQN actor (AutoRegressiveStochasticActor, networks.py:144-268) or the IQN actor
 * (StochasticActor, networks.py:271-323).  With the IQN actor the first eight actor slots hold, in tf.Module
 * order, merge_embedding_layer (D,200)+(200), noise_embedding_layer (64,d)+(d), output_action_layer (200,A)+(A),
 * state_embedding_layer (S,D)+(D) with d = 400 / A, D = d*A; the remaining five actor slots are empty. */
enum gac_actor_kind { GAC_ACTOR_AIQN = 0, GAC_ACTOR_IQN = 1 };
enum gac_iqn_var { GAC_I_WM = 0, GAC_I_BM, GAC_I_WN, GAC_I_BN, GAC_I_WO, GAC_I_BO, GAC_I_WS, GAC_I_BS };

/* Arena layout for (S, A).  offsets[v] = first float of variable v, rows/cols = its shape
 * (cols = 1 for vectors... biases are (n,) => rows=n, cols=1; gb is (2,1200)).
 * offsets[GAC_NVARS] = arena length in floats (each variable starts on a 128-byte boundary;
 * padding floats are zero and stay zero under Adam/Polyak).
 * net_begin[n]..net_begin[n+1] is the float range of network n. */
typedef struct gac_layout {
  int64_t offsets[GAC_NVARS + 1];
  int32_t rows[GAC_NVARS];
  int32_t cols[GAC_NVARS];
  int64_t net_begin[GAC_NNETS + 1];
} gac_layout_t;

int gac_layout(int32_t S, int32_t A, gac_layout_t* out /* host */);            /* AIQN actor */
int gac_layout_ex(int32_t S, int32_t A, int32_t actor_kind, gac_layout_t* out /* host */);

/* Static configuration of one agent replica (one per GPU). */
typedef struct gac_config {
  int32_t S, A;          /* state_dim, action_dim                       agent.py:49-50 */
  int32_t B;             /* states per step on THIS rank (batch_size)    agent.py:57    */
  int32_t K;             /* action_samples                               agent.py:53    */
  int32_t chunk_rows;    /* row chunk for the actor fwd/bwd workspace (0 = min(2*B*K, 65536)) */
  int32_t engine;        /* 0 = tcgen05 3xTF32 GEMMs where available, 1 = SIMT fp32 everywhere */
  int32_t actor_kind;    /* enum gac_actor_kind                          agent.py:65-70 */
} gac_config_t;

typedef struct gac_hparams {
  float gamma;           /* agent.py:52  */
  float rho;             /* Polyak rate `tau`, agent.py:56 */
  float q_normalization; /* agent.py:61  */
  float beta;            /* agent.py:55  */
  int32_t mode;          /* 0 = linear, 1 = boltzmann (networks.py:90-95) */
  float lr, beta1, beta2, eps; /* Keras Adam(1e-4): networks.py:77,381-382,439 */
} gac_hparams_t;

typedef struct gac_ctx gac_ctx_t;

/* Workspace sizing + context creation.  The context is a HOST object that only partitions
 * the caller-provided device workspace; it owns no device memory. */
size_t gac_workspace_bytes(const gac_config_t* cfg);
int gac_create(const gac_config_t* cfg, void* workspace, size_t workspace_bytes, void* stream,
               gac_ctx_t** out);
void gac_destroy(gac_ctx_t* ctx);
/* Re-bind the context to another stream (a cudaStream_t passed as void*): every later call is ordered on it.  The caller
 * orders the two streams (an event wait) if work issued on the old one may still touch the workspace.  Needed by callers that
 * change the current stream after construction, e.g. to capture a whole train step into a CUDA graph. */
int gac_set_stream(gac_ctx_t* ctx, void* stream);
const char* gac_last_error(void);
int gac_abi_version(void);

/* ---- helpers.update (GAC/helpers.py:75-86): t <- t*(1-rho) + s*rho, 128-bit streaming ---- */
int gac_polyak(float* target, const float* source, int64_t n, float rho, void* stream);

/* ---- Keras Adam x4 (networks.py:141,421,426,496) fused with the three Polyak updates
 * (agent.py:166-168) over the flat arenas.  adam_t: device int32[GAC_NNETS] step counters
 * (incremented here).  grad_scale: device float[GAC_NNETS] multipliers applied to each net's
 * gradient, skip: device int32[GAC_NNETS], non-zero = no Adam for that net this step
 * (agent.py:158: actor skipped when no positive-advantage row); Polyak always runs. ---- */
int gac_adam_polyak(gac_ctx_t* ctx, float* online, float* target, float* adam_m, float* adam_v, const float* grad,
                    const gac_hparams_t* hp /* host */, int32_t* adam_t, const float* grad_scale,
                    const int32_t* skip);

/* ---- AutoRegressiveStochasticActor.__call__ sampling mode (networks.py:194-225).
 * states_u: (n_states, S) unique states; state_idx: int32 (rows) row -> state;
 * taus: (rows, A); actions out: (rows, A).  actor: arena base of the actor variables. ---- */
int gac_aiqn_sample(gac_ctx_t* ctx, const float* actor, const float* states_u, int32_t n_states,
                    const int32_t* state_idx, const float* taus, int32_t rows, float* actions);

/* ---- one step of the actor's GRU (tf.keras.layers.GRU(400) cell, reset_after=True, gate order [z|r|h];
 * networks.py:166 / the per-dimension loop 205-225 and 258-262):
 *   h_out = GRU(x = [state_embedding(states_u)[state_idx] | action_emb], h_prev)
 * action_emb, h_prev, h_out: (rows, 400); h_prev == NULL means h_prev = 0 (first action dimension).
 * rows <= 2*B*K.  On the tcgen05 engine this is ONE kernel: both input projections, the gate math and the
 * state update (csrc/gru_step_tc.cuh).  flags:
 *   GAC_GRU_REUSE          states_u and the actor's weights are unchanged since the previous call on this context
 *                          (skips the per-state projection and the weight packing);
 *   GAC_GRU_ACTOR_INPUTS   action_emb is the actor's own embedding LeakyReLU(cos(.) wa + ba) and |h_prev| < 1, i.e.
 *                          both are bounded by the weights: allows the 3xFP16 arithmetic variant (same accuracy,
 *                          half the tensor-core instructions).  Without it the 3xTF32 variant runs, which has
 *                          fp32's exponent range.
 *   GAC_GRU_SAME_INPUTS    action_emb and h_prev hold the same values as in the previous call on this context (their
 *                          pre-split fp16 operand images are reused instead of rebuilt; only meaningful with ACTOR_INPUTS). ---- */
enum { GAC_GRU_REUSE = 1, GAC_GRU_ACTOR_INPUTS = 2, GAC_GRU_SAME_INPUTS = 4 };
int gac_gru_step(gac_ctx_t* ctx, const float* actor, const float* states_u, int32_t n_states, const int32_t* state_idx,
                 const float* action_emb, const float* h_prev, int32_t rows, float* h_out, int32_t flags);

/* ---- StochasticActor.__call__ (IQN, networks.py:304-323); same argument meaning as gac_aiqn_sample.
 * Context must have been created with actor_kind = GAC_ACTOR_IQN. ---- */
int gac_iqn_forward(gac_ctx_t* ctx, const float* actor, const float* states_u, int32_t n_states,
                    const int32_t* state_idx, const float* taus, int32_t rows, float* actions);
/* ---- the same forward keeping the activations for gac_iqn_bwd (IQNActor.train, networks.py:137-140) ---- */
int gac_iqn_train_fwd(gac_ctx_t* ctx, const float* actor, const float* states_u, int32_t n_states,
                      const int32_t* state_idx, const float* taus, int32_t max_rows, const int32_t* d_rows,
                      float* pred);
int gac_iqn_bwd(gac_ctx_t* ctx, const float* actor, const float* dpred, float* grad_actor, int32_t accumulate);

/* ---- Critic.__call__ (networks.py:384-388): q = min(fnn1, fnn2)(concat(s, a)) ---- */
int gac_critic_eval(gac_ctx_t* ctx, const float* fnn1, const float* fnn2, const float* states_u,
                    const int32_t* state_idx, const float* actions, int32_t rows, float* q);

/* ---- Value.__call__ (networks.py:441-442) on `rows` states ---- */
int gac_value_eval(gac_ctx_t* ctx, const float* value, const float* states, int32_t rows, float* v);

/* ---- tf.where(q > v) + gather_nd (agent.py:207-215).  n candidates; writes ascending int64
 * indices, exclusive positions, M (int32) and sum of advantages (fp32, deterministic order).
 * Optional gathers (may be NULL): adv_out[j] = q-v, actions_out (M,A) from actions (n,A),
 * sidx_out[j] = state_idx[idx[j]]. ---- */
int gac_advantage_compact(gac_ctx_t* ctx, const float* q, const float* v, int32_t n, int64_t* idx_out,
                          int32_t* pos_out, int32_t* m_out, float* sum_adv_out, float* adv_out,
                          const float* actions, int32_t A, float* actions_out,
                          const int32_t* state_idx, int32_t* sidx_out);

/* ---- _supervised_forward (networks.py:227-268) on rows given as (unique states, state_idx);
 * teacher actions (rows, A); pred out (rows, A).  d_rows: device int32 row count (<= max_rows)
 * or NULL for rows = max_rows.  Saves activations for gac_aiqn_bwd in the context. ---- */
int gac_aiqn_supervised_fwd(gac_ctx_t* ctx, const float* actor, const float* states_u, int32_t n_states,
                            const int32_t* state_idx, const float* taus, const float* teacher,
                            int32_t max_rows, const int32_t* d_rows, float* pred);

/* ---- IQNActor.huber_quantile_loss + d/dpred (networks.py:113-120, App. A.6).
 * weights (rows) multiply elementwise (pass adv for the un-normalised linear density);
 * inv_norm scales both loss and gradient (1/(M*A) or 1).  loss_out += sum. ---- */
int gac_qhuber_loss_bwd(gac_ctx_t* ctx, const float* pred, const float* target, const float* taus,
                        const float* weights, int32_t max_rows, const int32_t* d_rows, int32_t A,
                        float inv_norm, float* dpred, float* loss_out);

/* ---- BPTT through the activations saved by the last gac_aiqn_supervised_fwd
 * (tape.gradient at networks.py:140).  grad_actor: actor slice of the grad arena;
 * accumulate != 0 adds into it. ---- */
int gac_aiqn_bwd(gac_ctx_t* ctx, const float* actor, const float* dpred, float* grad_actor, int32_t accumulate);

/* ---- FNN TD regression: loss = sum_b (y_b - f(x_b))^2 and its gradient (networks.py:418-426,
 * 493-496; F5: gradient of the SUM).  x (rows, n_in); y (rows); grad_fnn: that net's slice of
 * the grad arena (overwritten); loss_out: device float (overwritten). ---- */
int gac_td_fwd_bwd(gac_ctx_t* ctx, const float* fnn, int32_t n_in, const float* x, const float* y,
                   int32_t rows, float* grad_fnn, float* loss_out);

/* ---- The whole GACAgent.train_one_step (agent.py:121-168), split at the data-parallel
 * exchange point: gac_step_grads fills `grad` (all four nets; the actor slice un-normalised)
 * and stats; the caller all-reduces (SUM) grad[0 .. n_total+GAC_STEP_NSTATS) across ranks when
 * world_size > 1; gac_step_apply normalises the actor gradient, runs Adam x4 and Polyak x3.
 * gac_step_grads runs part of its work (weight preparation, the TD updates, the actor's weight gradients) on a side stream
 * owned by the context, forked from and joined back into the context's stream by events: for the caller it stays one
 * stream-ordered call, and a stream capture of the context's stream captures both lanes (env GAC_NO_SIDE_STREAM=1: one lane). */
#define GAC_STEP_NSTATS 8
/* stats[] layout (floats appended to the grad arena at offset n_total):
 * 0 loss_fnn1, 1 loss_fnn2, 2 loss_value, 3 actor loss (un-normalised sum), 4 sum_adv, 5 M,
 * 6,7 reserved */
typedef struct gac_step_io {
  float* online; float* target; float* adam_m; float* adam_v;
  float* grad;                  /* n_total + GAC_STEP_NSTATS floats */
  int32_t* adam_t;              /* int32[GAC_NNETS] */
  const float* s; const float* a; const float* r; const float* sp; const float* done; /* replay batch */
  const float* critic_u;        /* (B,A)   U[0,1)   networks.py:411 */
  const float* value_tau;       /* (P,A)   U[0,1)   networks.py:473, state-major rows */
  const float* filter_tau;      /* (P,A)   U[0,1)   helpers.py:28,   sample-major rows */
  const float* filter_normal;   /* (P,A)   N(0,1)   agent.py:189 */
  const float* filter_rand;     /* (P,A)   U[-1,1)  agent.py:194 */
  const float* actor_tau;       /* (2P,A)  U[0,1)   networks.py:134, compacted row j reads row j */
  /* optional outputs (NULL to skip) */
  int64_t* sel_idx;             /* (2P) ascending selected candidate indices, first M valid */
  float* q_out; float* v_out;   /* (2P) candidate q and v as used by the filter */
  float* cand_actions;          /* (2P, A) candidate actions */
  float* value_actions;         /* (P, A) actions sampled for Value.train */
} gac_step_io_t;

int gac_step_grads(gac_ctx_t* ctx, const gac_step_io_t* io /* host */, const gac_hparams_t* hp /* host */);
int gac_step_apply(gac_ctx_t* ctx, const gac_step_io_t* io /* host */, const gac_hparams_t* hp /* host */);

/* ---- ReplayBuffer.sample_batch (GAC/helpers.py:65-72) on a device-resident ring buffer: gathers the five
 * transition arrays for `n` indices (int32, drawn by the caller: the reference uses random.sample) in one
 * kernel.  obs1/obs2 (size,S), acts (size,A), rews/done (size,1) -> s/sp (n,S), a (n,A), r/it (n,1). ---- */
int gac_replay_gather(const float* obs1, const float* obs2, const float* acts, const float* rews, const float* done,
                      const int32_t* idx, int32_t n, int32_t S, int32_t A, float* s, float* sp, float* a, float* r,
                      float* it, void* stream);

/* ---- plain GEMM entry used by the tests and the micro-benchmarks:
 * C(M,N) = A(M,K) * B(K,N), engine 0 = tcgen05 3xTF32, 1 = SIMT fp32,
 * 2 = tcgen05 3xTF32 with B pre-packed as a weight matrix (bulk-copy path). ---- */
int gac_gemm(gac_ctx_t* ctx, int32_t engine, const float* A, const float* B, float* C, int32_t M, int32_t N,
             int32_t K, int32_t transA, int32_t transB);

/* number of kernel launches issued through this context since creation (bench gpu_launches) */
int64_t gac_launch_count(const gac_ctx_t* ctx);

#ifdef __cplusplus
}
#endif
#endif /* GAC_B200_H */
