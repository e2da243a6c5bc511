// libgac_b200.so -- C ABI + orchestration of the GAC train step on B200 (sm_100a).
// See include/gac_b200.h for the contract and the reference file:line each entry replaces.
#include <math.h>
#include <stdlib.h>
#include <new>

#include "gac_common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gru_step_tc.cuh"
#include "gru_step_v2.cuh"
#include "gru_step_v3.cuh"
#include "dot_gemm_tc.cuh"

constexpr int FNN_BN = 160, FNN_NT = 2;     // FNN tail: 300 hidden columns in two 160-wide tiles
constexpr int HEAD_BN = 208, HEAD_NT = 2;   // actor head: 400 hidden columns in two 208-wide tiles (dual accumulators fit TMEM)
#include "kernels_elem.cuh"
#include "emb_gemm_tc.cuh"

namespace gac {
thread_local char g_err[512] = {0};
}
constexpr int G2_MAX_GRID = 256;               // upper bound of the persistent GRU kernel's grid (SM count; 148 on B200)
using namespace gac;

// ------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------
static const int kAlign = 32;  // floats (128 B)
static int64_t align_up64(int64_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

static void fnn_shapes(int n_in, int rows[6], int cols[6]) {
  const int r[6] = {n_in, GAC_H1, GAC_H1, GAC_H2, GAC_H2, 1};
  const int c[6] = {GAC_H1, 1, GAC_H2, 1, 1, 1};
  for (int i = 0; i < 6; ++i) { rows[i] = r[i]; cols[i] = c[i]; }
}
// offsets of the six FNN variables relative to the FNN's first float
static void fnn_offsets(int n_in, int64_t off[7]) {
  int rows[6], cols[6];
  fnn_shapes(n_in, rows, cols);
  int64_t o = 0;
  for (int i = 0; i < 6; ++i) { off[i] = o; o = align_up64(o + (int64_t)rows[i] * cols[i]); }
  off[6] = o;
}

extern "C" int gac_layout(int32_t S, int32_t A, gac_layout_t* out) { return gac_layout_ex(S, A, GAC_ACTOR_AIQN, out); }

extern "C" int gac_layout_ex(int32_t S, int32_t A, int32_t actor_kind, gac_layout_t* out) {
  if (!out || S <= 0 || A <= 0 || A > GAC_E) return fail(GAC_ERR_INVALID, "gac_layout: bad arguments");
  if (actor_kind != GAC_ACTOR_AIQN && actor_kind != GAC_ACTOR_IQN) return fail(GAC_ERR_INVALID, "gac_layout: unknown actor kind");
  int ar[13] = {GAC_NB, GAC_E, GAC_E, GAC_E, GAC_E, 1, GAC_NB, GAC_E, 2 * GAC_E, GAC_E, 2, S, GAC_E};
  int ac[13] = {GAC_E, 1, GAC_E, 1, 1, 1, GAC_E, 1, GAC_G, GAC_G, GAC_G, GAC_E, 1};
  if (actor_kind == GAC_ACTOR_IQN) {           // StochasticActor, networks.py:283-302
    const int d = GAC_E / A, D = d * A;
    const int ir[13] = {D, 200, GAC_NB, d, 200, A, S, D, 0, 0, 0, 0, 0};
    const int ic[13] = {200, 1, d, 1, A, 1, D, 1, 0, 0, 0, 0, 0};
    for (int i = 0; i < 13; ++i) { ar[i] = ir[i]; ac[i] = ic[i]; }
  }
  int64_t o = 0;
  int v = 0;
  out->net_begin[GAC_NET_ACTOR] = 0;
  for (int i = 0; i < 13; ++i, ++v) {
    out->offsets[v] = o; out->rows[v] = ar[i]; out->cols[v] = ac[i];
    o = align_up64(o + (int64_t)ar[i] * ac[i]);
  }
  const int nin[3] = {S + A, S + A, S};
  for (int f = 0; f < 3; ++f) {
    out->net_begin[GAC_NET_FNN1 + f] = o;
    int rows[6], cols[6];
    fnn_shapes(nin[f], rows, cols);
    for (int i = 0; i < 6; ++i, ++v) {
      out->offsets[v] = o; out->rows[v] = rows[i]; out->cols[v] = cols[i];
      o = align_up64(o + (int64_t)rows[i] * cols[i]);
    }
  }
  out->offsets[GAC_NVARS] = o;
  out->net_begin[GAC_NNETS] = o;
  return GAC_OK;
}

// ------------------------------------------------------------------------------------------
// context = partition of the caller's workspace
// ------------------------------------------------------------------------------------------
struct gac_ctx {
  gac_config_t cfg;
  gac_layout_t lay;
  cudaStream_t stream;
  int64_t launches;
  int P, Mc, nchunks, Rs, NUs, NUa, max_splits;
  const unsigned char* pk[16];   // packed pointer per slot (valid for the current call)
  int pk_bn[16];                 // tile width each slot was packed for (0 = one-tile-per-CTA kernel)
  // sampler / critic scratch (Rs rows per pass)
  float *se_u, *sx_u, *cosb, *ae, *mxa, *mh, *hA, *hB, *u, *y1;
  float *cx, *ch0a, *ch0b, *ch1a, *ch1b;
  // TD scratch (B rows)
  float *td_x, *td_h0, *td_h1, *td_pred, *td_dout, *td_da1, *td_da0, *td_y, *td_vnext, *td_vcrit;
  // step scratch
  float *taus2, *acts2, *cand, *q3, *vt, *v2, *adv_sel, *act_sel, *sum_adv, *blk_adv, *loss_part, *alpha, *gscale;
  int *sidx3, *pos, *sidx_sel, *blk_count, *m_dev, *rows_c, *skip;
  int64_t* idx_sel;
  // actor-train chunk (step-major [A][Mc][w])
  float *a_se_u, *a_sx_u, *a_dsx_u, *a_dse_u, *a_states; int* a_sidx;
  float *ca, *cn, *a_ae, *mx, *dmh, *zs, *rs, *cs, *mhh, *hall, *ne, *au, *ay1, *dy1p, *du, *dne, *dhu, *dhc0, *dhc1, *dop;
  float *pred_c, *dpred_c;
  const int* bwd_sidx; const float* bwd_states; int bwd_nstates; const int* bwd_drows;  // saved by supervised_fwd
  float *spart, *cpart, *fpart0, *fpart1, *rc_tmp;
  // IQN actor scratch (actor_kind == GAC_ACTOR_IQN): sampling (Rs rows) and training (Mc rows)
  int iq_d, iq_D;
  float *iq_se_u, *iq_cos, *iq_ne, *iq_me;
  float *it_x, *it_se, *it_cos, *it_ne, *it_merge, *it_me, *it_dop, *it_dme, *it_dmerge, *it_dne, *it_dse;
  int* it_live;   // [0] live rows of the chunk, [1] live rows * A
  unsigned char* pack[16]; size_t pack_bytes;   // pre-split / pre-swizzled weight tiles (tc::pack_b_kernel)
  int *seg_hist, *seg_perm;                     // per-block state histograms / rows grouped by state
  int* seg_info;                                // [unsorted | start[NU] | end[NU]] of the per-state segments (aiqn_bwd)
  // input-gradient GEMMs of the actor's backward on dot_gemm_tc.cuh (store mode): W1^T, U^T, kx_a^T tiles of both variants,
  // scales, and the running |x| maxima written by the kernels that produce their A operands ([0] dmx/dmh, [2] dy1p)
  unsigned char *w1f_pack32, *w1f_pack16; tc::DotScales* w1f_sc;   // training forward head GEMM (au * W1 + b1, LeakyReLU)
  unsigned char *dg_pack32[3], *dg_pack16[3]; tc::DotScales* dg_sc[3]; unsigned int* dg_amax; bool dg_fused;
  // 3xFP16 weight-gradient GEMMs of the actor (dW1, dU, dkx_a): scales per site, embedding bounds ([1] noise, [3] action)
  WgScales* wg_sc; unsigned int* wg_bound; bool wg_f16; unsigned char* wg_xpack; unsigned long long* dbg_buf;
  // cosine embeddings of the sampler as one kernel each (emb_gemm_tc.cuh): [0] action embedding, [1] noise embedding
  unsigned char* emb_pack_buf[4]; tc::EmbScales* emb_sc[4]; bool emb_fused, emb_train;   // [0], [1]: target actor (sampler); [2], [3]: online actor (training forward)
  // fused tails (400 -> 300 -> 1) of the twin critics in critic_eval
  unsigned char *cr_pack32[2], *cr_pack16[2]; tc::DotScales* cr_sc[2]; float* cr_part[2]; unsigned int* cr_bound;
  float* cr_sx0[2];                             // state half of the critics' first layer per unique state (n_states, 400)
  // fused actor head of the sampler (dot_gemm_tc.cuh): w1 tiles of both variants, scales, per-row tile partials
  unsigned char *head_pack32, *head_pack16; tc::DotScales* head_sc; float* head_part; bool head_fused;
  unsigned char* gru_pack16[2]; tc::GruScales* gru_sc[2];   // 3xFP16 variant: tiles + power-of-two scales (device)
  unsigned char* gru_pack[2];                   // fused GRU-step weights (tc::gru_pack_kernel): target / online actor
  const unsigned char *gru_t, *gru_o;           // valid for the current call; nullptr = unfused path
  // second-generation GRU step (gru_step_v2.cuh): weight tiles + scales per actor (target / online), a16 images of the
  // sampler's action embedding and of the two ping-pong GRU states
  unsigned char* g2_wpk[2]; tc::Gru2Scales* g2_sc[2]; unsigned char *ae_pk, *h_pk[2]; bool g2_on[2]; float* g2_xh;
  unsigned char* g3_wpk[2];                     // the same weights split per CTA of a pair (gru_step_v3.cuh, cta_group::2)
  // second lane of gac_step_grads: the three TD updates (B-row problems, ~65 small launches) run on a side stream beside the
  // sampler / critic evaluation / actor update; they touch only the td_* buffers and their own split-K / column-sum scratch
  cudaStream_t side; cudaEvent_t ev_fork, ev_q, ev_join, ev_cprep, ev_aprep; float *spart2, *cpart2;
  bool lanes_on; cudaStream_t main_stream; float *spart_main, *cpart_main; cudaEvent_t ev_w, ev_wjoin;   // lane state while gac_step_grads runs
  bool critic_prepped; const float* actor_prep_token;   // weight preparation already issued by gac_step_grads (on its side lane)
  unsigned char *a_ae_pk, *a_h_pk[2];           // training forward: a16 images of the (A * Mc)-row action embedding and of two ping-pong states
};

namespace {

struct Bump {
  char* base; size_t off;
  template <typename T> T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

int derive(gac_ctx* c) {
  const gac_config_t& f = c->cfg;
  if (f.S <= 0 || f.A <= 0 || f.B <= 0 || f.K <= 0) return fail(GAC_ERR_INVALID, "gac_config: S, A, B, K must be positive");
  const int64_t P = (int64_t)f.B * f.K;
  if (3 * P > (int64_t)1 << 30) return fail(GAC_ERR_INVALID, "gac_config: B*K too large");
  // the per-state gradient reduction keeps one shared-memory counter per unique state (seg_hist_kernel)
  if (f.B > 50 * 1024) return fail(GAC_ERR_INVALID, "gac_config: B (states per rank) must be <= 51200; shard the batch over more ranks");
  c->P = (int)P;
  int mc = f.chunk_rows > 0 ? f.chunk_rows : (int)(2 * P < 65536 ? 2 * P : 65536);
  c->Mc = round_up(mc, 128);
  c->nchunks = cdiv(2 * c->P, c->Mc);
  c->Rs = round_up((int)(2 * P < 131072 ? 2 * P : 131072), 128);
  c->NUs = c->Rs > f.B ? c->Rs : f.B;
  c->NUa = f.B > 1024 ? f.B : 1024;
  if (c->NUa > c->Mc && c->Mc >= f.B) c->NUa = c->Mc > f.B ? c->Mc : f.B;
  c->max_splits = 128;
  c->iq_d = GAC_E / f.A; c->iq_D = c->iq_d * f.A;
  return gac_layout_ex(f.S, f.A, f.actor_kind, &c->lay);
}

size_t partition(gac_ctx* c, void* ws) {
  Bump b{reinterpret_cast<char*>(ws), 0};
  const gac_config_t& f = c->cfg;
  const size_t Rs = c->Rs, B = f.B, P = c->P, A = f.A, S = f.S, Mc = c->Mc, AM = A * Mc;
  c->se_u = b.take<float>(c->NUs * GAC_E); c->sx_u = b.take<float>((size_t)c->NUs * GAC_G);
  c->cosb = b.take<float>(Rs * GAC_NB); c->ae = b.take<float>(Rs * GAC_E);
  c->mxa = b.take<float>(Rs * GAC_G); c->mh = b.take<float>(Rs * GAC_G);
  c->hA = b.take<float>(Rs * GAC_E); c->hB = b.take<float>(Rs * GAC_E);
  c->u = b.take<float>(Rs * GAC_E); c->y1 = b.take<float>(Rs * GAC_E);
  c->cx = b.take<float>(Rs * (S + A));
  c->ch0a = b.take<float>(Rs * GAC_H1); c->ch0b = b.take<float>(Rs * GAC_H1);
  c->ch1a = b.take<float>(Rs * GAC_H2); c->ch1b = b.take<float>(Rs * GAC_H2);
  c->td_x = b.take<float>(B * (S + A)); c->td_h0 = b.take<float>(B * GAC_H1); c->td_h1 = b.take<float>(B * GAC_H2);
  c->td_pred = b.take<float>(B); c->td_dout = b.take<float>(B); c->td_da1 = b.take<float>(B * GAC_H2);
  c->td_da0 = b.take<float>(B * GAC_H1); c->td_y = b.take<float>(B); c->td_vnext = b.take<float>(B);
  c->td_vcrit = b.take<float>(B);
  c->taus2 = b.take<float>(2 * P * A); c->acts2 = b.take<float>(2 * P * A); c->cand = b.take<float>(3 * P * A);
  c->q3 = b.take<float>(3 * P); c->vt = b.take<float>(B); c->v2 = b.take<float>(2 * P);
  c->adv_sel = b.take<float>(2 * P); c->act_sel = b.take<float>(2 * P * A); c->sum_adv = b.take<float>(1);
  const size_t nblk = (2 * P + CMP_BLOCK - 1) / CMP_BLOCK;
  c->blk_adv = b.take<float>(nblk); c->loss_part = b.take<float>((2 * P * A + 255) / 256 + 1);
  c->alpha = b.take<float>(GAC_NNETS); c->gscale = b.take<float>(GAC_NNETS);
  c->sidx3 = b.take<int>(3 * P); c->pos = b.take<int>(2 * P); c->sidx_sel = b.take<int>(2 * P);
  c->blk_count = b.take<int>(nblk); c->m_dev = b.take<int>(1); c->rows_c = b.take<int>(c->nchunks + 1);
  c->skip = b.take<int>(GAC_NNETS); c->idx_sel = b.take<int64_t>(2 * P);
  c->a_se_u = b.take<float>((size_t)c->NUa * GAC_E); c->a_sx_u = b.take<float>((size_t)c->NUa * GAC_G);
  c->a_dsx_u = b.take<float>((size_t)c->NUa * GAC_G); c->a_dse_u = b.take<float>((size_t)c->NUa * GAC_E);
  c->a_states = b.take<float>((size_t)c->NUa * S); c->a_sidx = b.take<int>(Mc); c->seg_info = b.take<int>(2 * (size_t)c->NUa + 2);
  c->seg_hist = b.take<int>((size_t)((Mc + SEG_BLOCK - 1) / SEG_BLOCK) * c->NUa); c->seg_perm = b.take<int>(Mc);
  c->ca = b.take<float>(AM * GAC_NB); c->cn = b.take<float>(AM * GAC_NB);
  c->a_ae = b.take<float>(AM * GAC_E); c->mx = b.take<float>(AM * GAC_G); c->dmh = b.take<float>(AM * GAC_G);
  c->zs = b.take<float>(AM * GAC_E); c->rs = b.take<float>(AM * GAC_E); c->cs = b.take<float>(AM * GAC_E);
  c->mhh = b.take<float>(AM * GAC_E); c->hall = b.take<float>((AM + Mc) * GAC_E);
  c->ne = b.take<float>(AM * GAC_E); c->au = b.take<float>(AM * GAC_E); c->ay1 = b.take<float>(AM * GAC_E);
  c->dy1p = b.take<float>(AM * GAC_E); c->du = b.take<float>(AM * GAC_E); c->dne = b.take<float>(AM * GAC_E);
  c->dhu = b.take<float>(AM * GAC_E); c->dhc0 = b.take<float>(Mc * GAC_E); c->dhc1 = b.take<float>(Mc * GAC_E);
  c->dop = b.take<float>(AM); c->pred_c = b.take<float>(Mc * A); c->dpred_c = b.take<float>(Mc * A);
  c->spart = b.take<float>((size_t)c->max_splits * GAC_E * GAC_G);
  c->spart2 = b.take<float>((size_t)c->max_splits * GAC_E * GAC_G);      // side lane: TD updates, then the actor's large weight gradients
  const size_t maxrows = AM > Rs ? AM : Rs;
  c->cpart = b.take<float>(((maxrows + COLSUM_ROWS - 1) / COLSUM_ROWS + 8192 / 16 + 1) * GAC_G);
  c->cpart2 = b.take<float>(((maxrows + COLSUM_ROWS - 1) / COLSUM_ROWS + 8192 / 16 + 1) * GAC_G);
  if (f.actor_kind == GAC_ACTOR_IQN) {
    const size_t D = c->iq_D;
    c->iq_se_u = b.take<float>((size_t)c->NUs * D); c->iq_cos = b.take<float>(Rs * A * GAC_NB); c->iq_ne = b.take<float>(Rs * D);
    c->iq_me = b.take<float>(Rs * 200);
    c->it_x = b.take<float>(Mc * S); c->it_se = b.take<float>(Mc * D); c->it_cos = b.take<float>(Mc * A * GAC_NB);
    c->it_ne = b.take<float>(Mc * D); c->it_merge = b.take<float>(Mc * D); c->it_me = b.take<float>(Mc * 200);
    c->it_dop = b.take<float>(Mc * A); c->it_dme = b.take<float>(Mc * 200); c->it_dmerge = b.take<float>(Mc * D);
    c->it_dne = b.take<float>(Mc * D); c->it_dse = b.take<float>(Mc * D); c->it_live = b.take<int>(4);
  }
  const size_t nrb = (AM + FB_ROWS - 1) / FB_ROWS + 1;
  c->fpart0 = b.take<float>(nrb * GAC_G); c->fpart1 = b.take<float>(nrb * GAC_G); c->rc_tmp = b.take<float>(64 * GAC_G);
  c->pack_bytes = tc_packed_bytes(GAC_E, GAC_G);
  if (tc_packed_bytes(GAC_G, GAC_E) > c->pack_bytes) c->pack_bytes = tc_packed_bytes(GAC_G, GAC_E);
  c->pack_bytes = (c->pack_bytes + 1023) & ~(size_t)1023;
  for (int i = 0; i < 16; ++i) c->pack[i] = b.take<unsigned char>(c->pack_bytes);
  for (int i = 0; i < 2; ++i) c->gru_pack[i] = b.take<unsigned char>(tc::GRU_PACK_BYTES);
  for (int i = 0; i < 2; ++i) c->gru_pack16[i] = b.take<unsigned char>(tc::GRU_PACK16_BYTES);
  for (int i = 0; i < 2; ++i) c->gru_sc[i] = b.take<tc::GruScales>(1);
  for (int i = 0; i < 2; ++i) { c->g2_wpk[i] = b.take<unsigned char>(tc::G2_WPK_BYTES); c->g2_sc[i] = b.take<tc::Gru2Scales>(1); }
  for (int i = 0; i < 2; ++i) c->g3_wpk[i] = b.take<unsigned char>(tc::G2_WPK_BYTES);
  c->ae_pk = b.take<unsigned char>(tc::a16_bytes((long long)Rs));
  for (int i = 0; i < 2; ++i) c->h_pk[i] = b.take<unsigned char>(tc::a16_bytes((long long)Rs));
  c->g2_xh = b.take<float>((size_t)G2_MAX_GRID * (tc::G2_XH_BYTES_PER_CTA / 4));
  c->a_ae_pk = b.take<unsigned char>(tc::a16_bytes((long long)AM));
  for (int i = 0; i < 2; ++i) c->a_h_pk[i] = b.take<unsigned char>(tc::a16_bytes((long long)Mc));
  c->head_pack32 = b.take<unsigned char>(tc::dot_pack_bytes(GAC_E, HEAD_BN, HEAD_NT, false));
  c->head_pack16 = b.take<unsigned char>(tc::dot_pack_bytes(GAC_E, HEAD_BN, HEAD_NT, true));
  c->head_sc = b.take<tc::DotScales>(1); c->head_part = b.take<float>((size_t)Rs * HEAD_NT);
  for (int i = 0; i < 3; ++i) {
    const int Kd = i == 0 ? GAC_E : GAC_G;
    c->dg_pack32[i] = b.take<unsigned char>(tc::dot_pack_bytes(Kd, HEAD_BN, HEAD_NT, false));
    c->dg_pack16[i] = b.take<unsigned char>(tc::dot_pack_bytes(Kd, HEAD_BN, HEAD_NT, true));
    c->dg_sc[i] = b.take<tc::DotScales>(1);
  }
  c->dg_amax = b.take<unsigned int>(4);
  c->w1f_pack32 = b.take<unsigned char>(tc::dot_pack_bytes(GAC_E, HEAD_BN, HEAD_NT, false));
  c->w1f_pack16 = b.take<unsigned char>(tc::dot_pack_bytes(GAC_E, HEAD_BN, HEAD_NT, true));
  c->w1f_sc = b.take<tc::DotScales>(1);
  for (int i = 0; i < 2; ++i) {
    c->cr_pack32[i] = b.take<unsigned char>(tc::dot_pack_bytes(GAC_H1, FNN_BN, FNN_NT, false));
    c->cr_pack16[i] = b.take<unsigned char>(tc::dot_pack_bytes(GAC_H1, FNN_BN, FNN_NT, true));
    c->cr_sc[i] = b.take<tc::DotScales>(1); c->cr_part[i] = b.take<float>((size_t)Rs * FNN_NT);
    c->cr_sx0[i] = b.take<float>((size_t)c->NUs * GAC_H1);
  }
  c->cr_bound = b.take<unsigned int>(4);
  for (int i = 0; i < 4; ++i) { c->emb_pack_buf[i] = b.take<unsigned char>(2 * tc::EMB_W_TILE); c->emb_sc[i] = b.take<tc::EmbScales>(1); }
  c->wg_sc = b.take<WgScales>(4); c->wg_bound = b.take<unsigned int>(4);
  c->wg_xpack = b.take<unsigned char>(wg_pack_x_bytes((long long)AM, GAC_E));
  c->dbg_buf = b.take<unsigned long long>(8 + 8 * 4096);   // per-CTA phase stamps of one GEMM launch (GAC_WGRAD_TIMELINE)
  return (b.off + 255) & ~(size_t)255;
}

#define LAUNCHED(ctx, name) do { (ctx)->launches++; GAC_LAUNCH_CHECK(name); } while (0)

inline const float* var(const gac_ctx* c, const float* net_base, int net, int v) {
  return net_base + (c->lay.offsets[v] - c->lay.net_begin[net]);
}
inline float* var(const gac_ctx* c, float* net_base, int net, int v) {
  return net_base + (c->lay.offsets[v] - c->lay.net_begin[net]);
}

// Pre-split + pre-swizzle a weight matrix for the tensor-core engine (nullptr on the SIMT engine).
enum { PK_T_WA = 0, PK_T_KXA, PK_T_U, PK_T_WN, PK_T_W1, PK_O_WA, PK_O_KXA, PK_O_U, PK_O_WN, PK_O_W1, PK_O_W1T, PK_O_UT, PK_O_KXAT,
       PK_C_F1, PK_C_F2 };
// rows = the smallest row count of the GEMMs that will use this image: large problems use the
// cta_group::2 persistent kernel (128-wide tiles), small ones the one-tile-per-CTA kernel.
const unsigned char* pack_w(gac_ctx* c, int slot, const float* W, int ld, int K, int N, int transB, int rows) {
  c->pk_bn[slot] = 0;
  if (c->cfg.engine != 0 || K < 32 || N < 32 || rows < 64 || tc_packed_bytes(K, N) > c->pack_bytes) return nullptr;   // rows < 64: SIMT path
  const int bn = (tc_pair_enabled() && rows >= 1024) ? tc::pair_bn(N) : 0;
  if (tc_pack_b(W, ld, transB, K, N, c->pack[slot], c->stream, bn) != cudaSuccess) return nullptr;
  c->launches++;
  c->pk_bn[slot] = bn;
  return c->pack[slot];
}

// cta_group::2 variant of the GRU step kernel: opt-in (GAC_GRU_PAIR=1).  First measured 205 vs 179 us per 32768-row step
// (profiles/r2d_gru_pair_probe.txt): every relayed stage paid a MEMBAR.ALL.GPU in front of its remote mbarrier arrive.  With
// default-semantics arrives, the instruction-group MMA issue and six stages it is level with the v2 kernel (163.5 vs 166.3 us
// in tools/gru2_probe.py, 162.1 vs 162.0 us in bench.py's loop, profiles/r2_pair_relay_probe.txt) -- both sit at the same
// power-capped wall -- so the single-CTA kernel, which every parity test of the round ran on, stays the default.
inline bool gru3_enabled() {
  static const bool pair = getenv("GAC_GRU_PAIR") != nullptr;
  return pair;
}
// weights of the fused GRU step (gru_step_tc.cuh); nullptr = not applicable, use the GEMM + gates kernels
// need_v1: the caller also runs round 1's kernel on these weights (gac_gru_step with operands it cannot bound); the sampler
// and the training forward only need it when the v2 kernel is switched off
const unsigned char* pack_gru(gac_ctx* c, int which, const float* actor, int rows, bool need_v1 = false) {
  static const bool off = getenv("GAC_NO_FUSED_GRU") != nullptr;
  static const bool tf32_only = getenv("GAC_GRU_TF32") != nullptr;   // A/B switch: skip the 3xFP16 variant
  static const bool v1_env = getenv("GAC_GRU_V1") != nullptr;
  if (off || c->cfg.engine != 0 || rows < 64) return nullptr;
  if (!v1_env && !tf32_only && !need_v1) {
    // the v2 kernel alone (it has its own fp32 path for weights that admit no fp16 scaling)
    c->g2_on[which] = gru2_pack(var(c, actor, 0, GAC_A_KX) + (size_t)GAC_E * GAC_G, var(c, actor, 0, GAC_A_U), var(c, actor, 0, GAC_A_WA),
                                var(c, actor, 0, GAC_A_BA), c->g2_wpk[which], c->g2_sc[which], c->stream) == cudaSuccess;
    if (c->g2_on[which]) {
      c->launches += 3;
      if (gru3_enabled()) {
        if (gru3_pack(var(c, actor, 0, GAC_A_KX) + (size_t)GAC_E * GAC_G, var(c, actor, 0, GAC_A_U), c->g3_wpk[which], c->g2_sc[which], c->stream) != cudaSuccess)
          c->g2_on[which] = false;
        else
          c->launches++;
      }
    }
    if (c->g2_on[which]) return c->g2_wpk[which];
  }
  if (gru_pack(var(c, actor, 0, GAC_A_KX) + (size_t)GAC_E * GAC_G, var(c, actor, 0, GAC_A_U), var(c, actor, 0, GAC_A_WA), var(c, actor, 0, GAC_A_BA),
               c->gru_pack[which], c->gru_pack16[which], tf32_only ? nullptr : c->gru_sc[which], c->stream) != cudaSuccess)
    return nullptr;
  c->launches += tf32_only ? 1 : 4;
  // second-generation kernel (persistent, bulk-copied pre-split A, double-buffered accumulators); GAC_GRU_V1=1 keeps round
  // 1's kernel for everything
  static const bool v1_only = getenv("GAC_GRU_V1") != nullptr;
  c->g2_on[which] = !v1_only && !tf32_only;
  if (c->g2_on[which]) {
    if (gru2_pack(var(c, actor, 0, GAC_A_KX) + (size_t)GAC_E * GAC_G, var(c, actor, 0, GAC_A_U), var(c, actor, 0, GAC_A_WA), var(c, actor, 0, GAC_A_BA),
                  c->g2_wpk[which], c->g2_sc[which], c->stream) != cudaSuccess)
      c->g2_on[which] = false;
    else
      c->launches += 3;
    if (c->g2_on[which] && gru3_enabled()) {
      if (gru3_pack(var(c, actor, 0, GAC_A_KX) + (size_t)GAC_E * GAC_G, var(c, actor, 0, GAC_A_U), c->g3_wpk[which], c->g2_sc[which], c->stream) != cudaSuccess)
        c->g2_on[which] = false;
      else
        c->launches++;
    }
  }
  return c->gru_pack[which];
}
// the v2 kernel on a16 operand images; it exits at once when its scales are not usable, and round 1's kernel (launched
// right after with skip_if = &scales->ok) then does the step from the fp32 operands
int run_gru_step_v2(gac_ctx* c, int which, const float* actor, const float* sx, const int* sidx, const unsigned char* ae_pk, const float* ae32,
                    const unsigned char* h_pk, const float* hprev, float* hout, unsigned char* hout_pk, int rows, const int* dyn,
                    float* z = nullptr, float* r = nullptr, float* cc = nullptr, float* mhh = nullptr) {
  static const bool timeline = getenv("GAC_GRU_TIMELINE") != nullptr;
  tc::Gru2Args a;
  const bool pair = gru3_enabled();
  a.rows = rows; a.dyn_rows = dyn; a.ae_pk = ae_pk; a.h_pk = hprev ? h_pk : nullptr; a.hprev = hprev; a.wpk = pair ? c->g3_wpk[which] : c->g2_wpk[which]; a.sc = c->g2_sc[which];
  a.sx = sx; a.sidx = sidx; a.b1 = var(c, actor, 0, GAC_A_GB) + GAC_G; a.hout = hout; a.hout_pk = hout_pk;
  a.z = z; a.r = r; a.c = cc; a.mhh = mhh; a.xh_scratch = c->g2_xh;
  a.ae32 = ae32; a.kxa32 = var(c, actor, 0, GAC_A_KX) + (size_t)GAC_E * GAC_G; a.U32 = var(c, actor, 0, GAC_A_U);
  static const int probe = getenv("GAC_G2_PROBE") ? atoi(getenv("GAC_G2_PROBE")) : 0;
  a.probe = probe;
  a.dbg = timeline ? reinterpret_cast<unsigned long long*>(c->spart) : nullptr;
  GAC_CUDA(pair ? launch_gru_step_v3(a, c->stream) : launch_gru_step_v2(a, c->stream));
  c->launches++;
  return GAC_OK;
}
int run_gru_step(gac_ctx* c, int which, const float* actor, const float* sx, const int* sidx, const float* ae,
                 const float* hprev, float* hout, int rows, const int* dyn, float* z, float* r, float* cc, float* mhh, bool bounded = true,
                 const int* skip_if = nullptr) {
  static const bool timeline = getenv("GAC_GRU_TIMELINE") != nullptr;
  static const bool tf32_only = getenv("GAC_GRU_TF32") != nullptr;
  tc::GruStepArgs a;
  a.dbg = timeline ? reinterpret_cast<unsigned long long*>(c->spart) : nullptr;   // scratch: idle during the sampler
  a.sc = (tf32_only || !bounded) ? nullptr : c->gru_sc[which];   // 3xFP16 needs operands bounded by the weights
  a.rows = rows; a.dyn_rows = dyn; a.ae = ae; a.hprev = hprev; a.wpk = c->gru_pack[which]; a.sx = sx; a.sidx = sidx;
  a.b1 = var(c, actor, 0, GAC_A_GB) + GAC_G; a.hout = hout; a.z = z; a.r = r; a.c = cc; a.mhh = mhh;
  a.wpk16 = tf32_only ? nullptr : c->gru_pack16[which];
  a.skip_if = skip_if;
  GAC_CUDA(launch_gru_step(a, c->stream));
  c->launches++;
  return GAC_OK;
}

// GEMM dispatch: tcgen05 3xTF32 engine where the shape qualifies, else SIMT; split-K for the
// weight-gradient contractions (reduction over activation rows), reduced deterministically.
int run_gemm(gac_ctx* c, GemmArgs g) {
  cudaStream_t st = c->stream;
  if (g.M <= 0 || g.N <= 0) return GAC_OK;
  const WgScales* f16_in = g.f16;
  const unsigned char* bpk_in = g.Bpk16;
  g.f16 = nullptr; g.Bpk16 = nullptr;
  const bool tc = c->cfg.engine == 0 && tc_gemm_supported(g);
  if (g.valid_on_k || g.K >= 8192) {
    // weight-gradient shape: small output, long reduction over activation rows => split-K.
    // On the tensor-core engine compute the transposed product (dY^T X) so that the LONG weight
    // dimension fills the 128-row MMA tile and the TMEM-lane-per-thread epilogue stores coalesce.
    if (tc && g.transA && !g.transB && !g.transC && g.ldc == g.N && g.N >= g.M) {
      GemmArgs w = g;
      w.M = g.N; w.N = g.M;
      w.A = g.B; w.lda = g.ldb; w.transA = 1;      // A'(m', k) = dY[k][m']
      w.B = g.A; w.ldb = g.lda; w.transB = 0;      // B'(k, n') = X[k][n']
      w.transC = 1;                                // C'(m', n') -> dW[n'][m'],  ldc = original N
      w.f16 = f16_in; w.Bpk16 = bpk_in;            // the 3xFP16 scales / packed X are laid out for this swapped product (A' = dY, B' = X)
      g = w;
    }
    const int tiles = tc ? cdiv(g.M, tc::BM) * cdiv(g.N, tc::plan_for(g.N).bn) : cdiv(g.M, 128) * cdiv(g.N, 128);
    int splits = (2 * 148) / tiles;
    // tcgen05 truncates toward zero when it adds into the fp32 accumulator (measured: relative
    // shrink ~2.0e-9 per reduction element, tools/gemm_accuracy.py), so bound the chain that one
    // TMEM accumulator sees to 4096 reduction rows (<= 8e-6); the partials are summed in fp32 (round to nearest)
    if (tc && splits < cdiv(g.K, 4096)) splits = cdiv(g.K, 4096);
    static const bool old_splits = getenv("GAC_WG_OLD_SPLITS") != nullptr;   // A/B switch
    if (tc && g.valid_on_k && !old_splits) {
      // Whole waves: tiles x splits a multiple of the SM count (the kernel keeps one CTA per SM), e.g. 20 tiles x 37 splits =
      // 5 x 148 instead of 20 x 48 = 6.49 waves, with fewer partial tiles to write and fold.  Dead rows are skipped, so a split
      // rarely sees the 4096 rows the rule above allows; when every row is live the chain grows to <= 6144 rows (bias <= 1.2e-5).
      static int nsm = 0;
      if (!nsm) { int dev = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) nsm = 148; }
      for (int s2 = cdiv(g.K, 6144); s2 <= c->max_splits && s2 <= splits; ++s2)
        if (s2 >= 2 && (tiles * s2) % nsm == 0) { splits = s2; break; }
    }
    if (splits > c->max_splits) splits = c->max_splits;
    if (splits > g.K / 512) splits = g.K / 512;
    if ((size_t)g.M * g.N > (size_t)GAC_E * GAC_G) splits = 1;
    if (splits > 1) {
      const bool dense = g.transC ? (g.ldc == g.M) : (g.ldc == g.N);
      if (!dense) return fail(GAC_ERR_INVALID, "split-K destination must be dense");
      GemmArgs p = g;
      p.C = c->spart; p.splitk = splits; p.split_stride = (long long)g.M * g.N;
      p.bias = nullptr; p.add = nullptr; p.mul = nullptr; p.gate = nullptr; p.act = ACT_NONE; p.accumulate = 0;
      if (tc) GAC_CUDA(launch_gemm_tc(p, st)); else GAC_CUDA(launch_gemm_simt(p, st));
      c->launches++;
      const long long n = (long long)g.M * g.N;
      reduce_splits_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(c->spart, n, splits, g.C, n, g.accumulate);
      LAUNCHED(c, "reduce_splits");
      return GAC_OK;
    }
  }
  // Small problems (the B-row TD nets, the per-state projections): a handful of tiles would walk K serially on a few
  // SMs -- latency bound, 17-70 us for a few MFLOP.  Spread K over the chip and apply the epilogue in the reduction.
  static const bool no_small_split = getenv("GAC_NO_SMALL_SPLITK") != nullptr;
  if (!no_small_split && !g.valid_on_k && !g.dyn_rows && g.K >= 128 && g.M > 8 && (size_t)g.M * g.N <= (size_t)GAC_E * GAC_G && g.act != ACT_TANH) {
    const int tiles = tc ? cdiv(g.M, tc::BM) * cdiv(g.N, tc::plan_for(g.N).bn) : cdiv(g.M, 128) * cdiv(g.N, 128);
    int splits = cdiv(g.K, 64);                                // 64-wide K slices (two tcgen05 K blocks)
    if (splits > 148 / tiles) splits = 148 / tiles;
    if (splits > c->max_splits) splits = c->max_splits;
    if (tiles <= 24 && splits >= 2) {
      GemmArgs p = g;
      p.C = c->spart; p.ldc = g.N; p.transC = 0; p.splitk = splits; p.split_stride = (long long)g.M * g.N;
      p.bias = nullptr; p.add = nullptr; p.mul = nullptr; p.gate = nullptr; p.act = ACT_NONE; p.accumulate = 0;
      p.Bpacked = nullptr; p.Bpacked_bn = 0;                   // K-split tiles read B through the register path
      if (tc) GAC_CUDA(launch_gemm_tc(p, st)); else GAC_CUDA(launch_gemm_simt(p, st));
      c->launches++;
      const long long n = (long long)g.M * g.N;
      reduce_splits_epi_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(c->spart, splits, g);
      LAUNCHED(c, "reduce_splits_epi");
      return GAC_OK;
    }
  }
  if (tc) GAC_CUDA(launch_gemm_tc(g, st)); else GAC_CUDA(launch_gemm_simt(g, st));
  c->launches++;
  return GAC_OK;
}

// dst[n] (+)= sum of nsplit partial rows of N columns, fixed order; many partial rows are folded in two stages (main lane only:
// one scratch buffer)
int run_reduce_cols(gac_ctx* c, const float* part, int nsplit, int N, float* dst, int accumulate) {
  if (nsplit >= 512) {
    const int slices = nsplit >= 2048 ? 32 : 16;
    reduce_cols_kernel<<<dim3(cdiv(N, 32), slices), dim3(32, RC_LANES), 0, c->stream>>>(part, nsplit, N, c->rc_tmp, 0);
    LAUNCHED(c, "reduce_cols_slices");
    reduce_cols_kernel<<<cdiv(N, 32), dim3(32, RC_LANES), 0, c->stream>>>(c->rc_tmp, slices, N, dst, accumulate);
    LAUNCHED(c, "reduce_cols");
    return GAC_OK;
  }
  reduce_cols_kernel<<<cdiv(N, 32), dim3(32, RC_LANES), 0, c->stream>>>(part, nsplit, N, dst, accumulate);
  LAUNCHED(c, "reduce_cols");
  return GAC_OK;
}

// dst[n] (+)= sum_r w[r] * X[r][n] over live rows
int run_colsum(gac_ctx* c, const float* X, int ldx, const float* w, int N, long long rows, RowValid rv, float* dst, int accumulate) {
  const int rpb = rows <= 8192 ? 16 : COLSUM_ROWS;
  const int nblk = (int)cdiv64(rows, rpb);
  colsum_kernel<<<dim3(cdiv(N, 128), nblk), 128, 0, c->stream>>>(X, ldx, w, N, rows, rv, c->cpart, rpb);
  LAUNCHED(c, "colsum");
  reduce_cols_kernel<<<cdiv(N, 32), dim3(32, RC_LANES), 0, c->stream>>>(c->cpart, nblk, N, dst, accumulate);
  LAUNCHED(c, "colsum_reduce");
  return GAC_OK;
}

// FNN forward on `rows` rows of x (rows, n_in): h0, h1 saved, out (rows) = linear head
int fnn_forward(gac_ctx* c, const float* fnn, int n_in, const float* x, int rows, float* h0, float* h1, const unsigned char* w1_packed = nullptr, int w1_bn = 0) {
  int64_t off[7];
  fnn_offsets(n_in, off);
  GemmArgs g = gemm_args(rows, GAC_H1, n_in, x, n_in, fnn + off[0], GAC_H1, h0, GAC_H1);
  g.bias = fnn + off[1]; g.act = ACT_LRELU; g.alpha = 0.01f;
  GAC_TRY(run_gemm(c, g));
  g = gemm_args(rows, GAC_H2, GAC_H1, h0, GAC_H1, fnn + off[2], GAC_H2, h1, GAC_H2);
  g.bias = fnn + off[3]; g.act = ACT_LRELU; g.alpha = 0.01f; g.Bpacked = w1_packed; g.Bpacked_bn = w1_bn;
  GAC_TRY(run_gemm(c, g));
  return GAC_OK;
}

int value_eval(gac_ctx* c, const float* value, const float* states, int rows, float* v, float* h0 = nullptr, float* h1 = nullptr) {
  const int S = c->cfg.S;
  if (!h0) { h0 = c->ch0a; h1 = c->ch1a; }                        // hidden-layer scratch (Rs rows); the side lane passes its own
  int64_t off[7];
  fnn_offsets(S, off);
  for (int r0 = 0; r0 < rows; r0 += c->Rs) {
    const int n = rows - r0 < c->Rs ? rows - r0 : c->Rs;
    GAC_TRY(fnn_forward(c, value, S, states + (size_t)r0 * S, n, h0, h1));
    rowdot_kernel<<<cdiv(n * 32, 256), 256, 0, c->stream>>>(h1, GAC_H2, value + off[4], value + off[5], nullptr, nullptr,
                                                           nullptr, GAC_H2, n, 0, v + r0, 1, nullptr, 0, RowValid{nullptr, 1});
    LAUNCHED(c, "rowdot");
  }
  return GAC_OK;
}

// Fused critic tail (dot_gemm_tc.cuh): layer 2 + LeakyReLU + output dot in one kernel, 3xFP16 with the bound of the first
// hidden layer derived from the weights and the batch's states (|action| <= 1).  Needs the number of unique states.
inline bool critic_fused(const gac_ctx* c, int rows, int n_states, const int* sidx) {
  static const bool off_env = getenv("GAC_NO_FUSED_CRITIC") != nullptr;
  return !off_env && c->cfg.engine == 0 && rows >= 1024 && n_states > 0 && n_states <= c->NUs && sidx != nullptr;
}
// everything of the fused critic evaluation that depends on the weights and the unique states only (on c->stream)
int critic_prep(gac_ctx* c, const float* f1, const float* f2, const float* states_u, int n_states) {
  const int S = c->cfg.S, A = c->cfg.A, W = S + A;
  cudaStream_t st = c->stream;
  int64_t off[7];
  fnn_offsets(W, off);
  const float* fn[2] = {f1, f2};
  GAC_CUDA(cudaMemsetAsync(c->cr_bound, 0, 4 * sizeof(unsigned int), st));
  for (int i = 0; i < 2; ++i) {
    fnn_bound_kernel<<<1, 512, W * sizeof(float), st>>>(states_u, n_states, S, W, fn[i] + off[0], fn[i] + off[1], GAC_H1, c->cr_bound + i);
    LAUNCHED(c, "fnn_bound");
    GAC_CUDA(dot_pack(fn[i] + off[2], GAC_H2, 0, GAC_H1, GAC_H2, FNN_BN, FNN_NT, nullptr, nullptr, 0, 0, c->cr_pack32[i], c->cr_pack16[i],
                      c->cr_sc[i], st));
    GAC_CUDA(dot_ascale(c->cr_sc[i], c->cr_bound + i, st));
    c->launches += 5;
    // state half of layer 1, once per unique state: sx0 = s_u W0[:S] + b0
    GemmArgs g0 = gemm_args(n_states, GAC_H1, S, states_u, S, fn[i] + off[0], GAC_H1, c->cr_sx0[i], GAC_H1);
    g0.bias = fn[i] + off[1];
    GAC_TRY(run_gemm(c, g0));
  }
  return GAC_OK;
}

int critic_eval(gac_ctx* c, const float* f1, const float* f2, const float* states_u, const int* sidx, const float* actions,
                int rows, float* q, int n_states = 0) {
  const int S = c->cfg.S, A = c->cfg.A, W = S + A;
  cudaStream_t st = c->stream;
  int64_t off[7];
  fnn_offsets(W, off);
  const bool fused = critic_fused(c, rows, n_states, sidx);
  const float* fn[2] = {f1, f2};
  const unsigned char *pk1 = nullptr, *pk2 = nullptr;
  if (fused) {
    if (c->critic_prepped) c->critic_prepped = false;             // gac_step_grads issued it on its side lane
    else GAC_TRY(critic_prep(c, f1, f2, states_u, n_states));
  } else if (rows >= 1024) {
    const int prow = rows < c->Rs ? rows : (rows % c->Rs ? rows % c->Rs : c->Rs);
    pk1 = pack_w(c, PK_C_F1, f1 + off[2], GAC_H2, GAC_H1, GAC_H2, 0, prow);
    pk2 = pack_w(c, PK_C_F2, f2 + off[2], GAC_H2, GAC_H1, GAC_H2, 0, prow);
  }
  for (int r0 = 0; r0 < rows; r0 += c->Rs) {
    const int n = rows - r0 < c->Rs ? rows - r0 : c->Rs;
    if (fused) {
      float* h0[2] = {c->ch0a, c->ch0b};
      for (int i = 0; i < 2; ++i) {
        // layer 1: gathered state half + A FMAs per output (no (S + A)-deep GEMM over every candidate row)
        fnn_l1_gather_kernel<<<(unsigned)cdiv64((long long)n * (GAC_H1 / 4), 256), 256, 0, st>>>(
            c->cr_sx0[i], sidx + r0, actions + (size_t)r0 * A, fn[i] + off[0] + (size_t)S * GAC_H1, A, GAC_H1, n, h0[i]);
        LAUNCHED(c, "fnn_l1_gather");
        tc::DotGemmArgs d;
        d.rows = n; d.dyn_rows = nullptr; d.seg_rows = 0; d.A = h0[i]; d.lda = GAC_H1; d.K = GAC_H1; d.N = GAC_H2; d.bn = FNN_BN; d.ntiles = FNN_NT;
        d.wpk = c->cr_pack32[i]; d.wpk16 = c->cr_pack16[i]; d.sc = c->cr_sc[i];
        d.bias = fn[i] + off[3]; d.w2 = fn[i] + off[4]; d.alpha = 0.01f; d.part = c->cr_part[i];
        d.mode = 0; d.C = nullptr; d.ldc = 0; d.gate = nullptr; d.ldgate = 0; d.gate_alpha = 0.f; d.accumulate = 0; d.act_lrelu = 0; d.dbg = nullptr;
        GAC_CUDA(launch_dot_gemm(d, st));
        c->launches++;
      }
      critic_finish_kernel<<<cdiv(n, 256), 256, 0, st>>>(c->cr_part[0], c->cr_part[1], FNN_NT, f1 + off[5], f2 + off[5], q + r0, n);
      LAUNCHED(c, "critic_finish");
      continue;
    }
    concat_gather_kernel<<<(unsigned)cdiv64((long long)n * W, 256), 256, 0, st>>>(
        states_u, sidx ? sidx + r0 : nullptr, actions + (size_t)r0 * A, S, A, n, c->cx);
    LAUNCHED(c, "concat_gather");
    GAC_TRY(fnn_forward(c, f1, W, c->cx, n, c->ch0a, c->ch1a, pk1, c->pk_bn[PK_C_F1]));
    GAC_TRY(fnn_forward(c, f2, W, c->cx, n, c->ch0b, c->ch1b, pk2, c->pk_bn[PK_C_F2]));
    rowdot_kernel<<<cdiv(n * 32, 256), 256, 0, st>>>(c->ch1a, GAC_H2, f1 + off[4], f1 + off[5], c->ch1b, f2 + off[4],
                                                     f2 + off[5], GAC_H2, n, 0, q + r0, 1, nullptr, 0, RowValid{nullptr, 1});
    LAUNCHED(c, "rowdot_min");
  }
  return GAC_OK;
}

// per-unique-state prep: se = lrelu_.2(lrelu_.01(s Ws + bs)); sx = se Kx[0:400] + b_in
int state_prep(gac_ctx* c, const float* actor, const float* states_u, int n_states, float* se_u, float* sx_u) {
  const int S = c->cfg.S;
  GemmArgs g = gemm_args(n_states, GAC_E, S, states_u, S, var(c, actor, 0, GAC_A_WS), GAC_E, se_u, GAC_E);
  g.bias = var(c, actor, 0, GAC_A_BS); g.act = ACT_LRELU2;
  GAC_TRY(run_gemm(c, g));
  g = gemm_args(n_states, GAC_G, GAC_E, se_u, GAC_E, var(c, actor, 0, GAC_A_KX), GAC_G, sx_u, GAC_G);
  g.bias = var(c, actor, 0, GAC_A_GB);
  GAC_TRY(run_gemm(c, g));
  return GAC_OK;
}

int aiqn_sample_rows(gac_ctx* c, const float* actor, const int* sidx, const float* taus, int rows, float* actions) {
  const int A = c->cfg.A;
  cudaStream_t st = c->stream;
  const RowValid all{nullptr, 1};
  float* hprev = c->hA; float* hcur = c->hB;
  unsigned char* pk_prev = c->h_pk[0]; unsigned char* pk_cur = c->h_pk[1];
  const bool v2 = c->gru_t && c->g2_on[0];
  const float* kxa = var(c, actor, 0, GAC_A_KX) + (size_t)GAC_E * GAC_G;
  for (int d = 0; d < A; ++d) {
    GemmArgs g;
    // cosine embedding (basis + Dense + activation [+ Hadamard]) as one kernel; x_r = x[r * A]
    auto embed = [&](int which, const float* x, const float* W, const float* bias, int act, float alpha, const float* mul, float* out) -> int {
      tc::EmbArgs e;
      e.rows = rows; e.dyn_rows = nullptr; e.seg_rows = 0; e.x = x; e.incx = A; e.xseg = 0; e.xoff = 0; e.wpk = c->emb_pack_buf[which]; e.sc = c->emb_sc[which];
      e.W = W; e.bias = bias; e.act = act; e.alpha = alpha; e.mul = mul; e.ldmul = GAC_E; e.C = out; e.ldc = GAC_E; e.C2 = nullptr;
      // the action embedding feeds only the GRU step: with the v2 kernel it is written as that kernel's A-operand image
      e.C16 = (which == 0 && v2) ? c->ae_pk : nullptr; e.c16_scale = &c->g2_sc[0]->s_ae; e.c16_ok = &c->g2_sc[0]->ok; e.c16_also = 0;
      GAC_CUDA(launch_emb_gemm(e, st));
      c->launches++;
      return GAC_OK;
    };
    if (c->emb_fused) {
      GAC_TRY(embed(0, d ? actions + d - 1 : nullptr, var(c, actor, 0, GAC_A_WA), var(c, actor, 0, GAC_A_BA), 1, 0.2f, nullptr, c->ae));
    } else {
      if (!(d && c->head_fused)) {   // the fused head's finish kernel already wrote cos(a_{d-1})
        cos_basis_kernel<<<cdiv(rows * (GAC_NB / 4), 256), 256, 0, st>>>(d ? actions + d - 1 : nullptr, A, c->cosb, rows, all);
        LAUNCHED(c, "cos_basis");
      }
      g = gemm_args(rows, GAC_E, GAC_NB, c->cosb, GAC_NB, var(c, actor, 0, GAC_A_WA), GAC_E, c->ae, GAC_E);
      g.bias = var(c, actor, 0, GAC_A_BA); g.act = ACT_LRELU; g.alpha = 0.2f; g.Bpacked = c->pk[PK_T_WA]; g.Bpacked_bn = c->pk_bn[PK_T_WA];
      GAC_TRY(run_gemm(c, g));
      if (v2) { GAC_CUDA(a16_pack(c->ae, GAC_E, rows, nullptr, 0, &c->g2_sc[0]->s_ae, c->ae_pk, st)); c->launches++; }
    }
    if (c->gru_t) {
      // both projections + gates in one kernel: mxa / mh never touch HBM
      // (the last dimension's state feeds no further GRU step: no operand image of it)
      if (v2) GAC_TRY(run_gru_step_v2(c, 0, actor, c->sx_u, sidx, c->ae_pk, c->ae, pk_prev, d ? hprev : nullptr, hcur, d + 1 < A ? pk_cur : nullptr, rows, nullptr));
      else GAC_TRY(run_gru_step(c, 0, actor, c->sx_u, sidx, c->ae, d ? hprev : nullptr, hcur, rows, nullptr, nullptr, nullptr, nullptr, nullptr, true, nullptr));
    } else {
      g = gemm_args(rows, GAC_G, GAC_E, c->ae, GAC_E, kxa, GAC_G, c->mxa, GAC_G);
      g.Bpacked = c->pk[PK_T_KXA]; g.Bpacked_bn = c->pk_bn[PK_T_KXA];
      GAC_TRY(run_gemm(c, g));
      if (d) {
        g = gemm_args(rows, GAC_G, GAC_E, hprev, GAC_E, var(c, actor, 0, GAC_A_U), GAC_G, c->mh, GAC_G);
        g.Bpacked = c->pk[PK_T_U]; g.Bpacked_bn = c->pk_bn[PK_T_U];
        GAC_TRY(run_gemm(c, g));
      }
      gru_gates_fwd_kernel<<<(unsigned)cdiv64((long long)rows * GAC_E, 256), 256, 0, st>>>(
          c->sx_u, sidx, c->mxa, d ? c->mh : nullptr, var(c, actor, 0, GAC_A_GB) + GAC_G, hprev, hcur, nullptr, nullptr, nullptr,
          nullptr, rows, all);
      LAUNCHED(c, "gru_gates_fwd");
    }
    if (c->emb_fused) {
      GAC_TRY(embed(1, taus + d, var(c, actor, 0, GAC_A_WN), var(c, actor, 0, GAC_A_BN), 0, 0.f, hcur, c->u));
    } else {
      cos_basis_kernel<<<cdiv(rows * (GAC_NB / 4), 256), 256, 0, st>>>(taus + d, A, c->cosb, rows, all);
      LAUNCHED(c, "cos_basis");
      g = gemm_args(rows, GAC_E, GAC_NB, c->cosb, GAC_NB, var(c, actor, 0, GAC_A_WN), GAC_E, c->u, GAC_E);
      g.bias = var(c, actor, 0, GAC_A_BN); g.mul = hcur; g.ldmul = GAC_E; g.Bpacked = c->pk[PK_T_WN]; g.Bpacked_bn = c->pk_bn[PK_T_WN];
      GAC_TRY(run_gemm(c, g));
    }
    if (c->head_fused) {
      static const bool tf32_only = getenv("GAC_HEAD_TF32") != nullptr;
      tc::DotGemmArgs q;
      q.rows = rows; q.dyn_rows = nullptr; q.A = c->u; q.lda = GAC_E; q.K = GAC_E; q.N = GAC_E; q.bn = HEAD_BN; q.ntiles = HEAD_NT;
      q.wpk = c->head_pack32; q.wpk16 = c->head_pack16; q.sc = tf32_only ? nullptr : c->head_sc;
      q.bias = var(c, actor, 0, GAC_A_B1); q.w2 = var(c, actor, 0, GAC_A_W2); q.alpha = 0.01f; q.part = c->head_part;
      q.mode = 0; q.seg_rows = 0; q.C = nullptr; q.ldc = 0; q.gate = nullptr; q.ldgate = 0; q.gate_alpha = 0.f; q.accumulate = 0; q.act_lrelu = 0; q.dbg = nullptr;
      GAC_CUDA(launch_dot_gemm(q, st));
      c->launches++;
      head_finish_kernel<<<cdiv(rows * (GAC_NB / 4), 256), 256, 0, st>>>(c->head_part, HEAD_NT, var(c, actor, 0, GAC_A_B2), actions + d, A,
                                                                   (d + 1 < A && !c->emb_fused) ? c->cosb : nullptr, rows);
      LAUNCHED(c, "head_finish");
    } else {
      g = gemm_args(rows, GAC_E, GAC_E, c->u, GAC_E, var(c, actor, 0, GAC_A_W1), GAC_E, c->y1, GAC_E);
      g.bias = var(c, actor, 0, GAC_A_B1); g.act = ACT_LRELU; g.alpha = 0.01f; g.Bpacked = c->pk[PK_T_W1]; g.Bpacked_bn = c->pk_bn[PK_T_W1];
      GAC_TRY(run_gemm(c, g));
      rowdot_kernel<<<cdiv(rows * 32, 256), 256, 0, st>>>(c->y1, GAC_E, var(c, actor, 0, GAC_A_W2), var(c, actor, 0, GAC_A_B2),
                                                         nullptr, nullptr, nullptr, GAC_E, rows, 1, actions + d, A, nullptr, 0, all);
      LAUNCHED(c, "rowdot_tanh");
    }
    float* t = hprev; hprev = hcur; hcur = t;
    unsigned char* tp = pk_prev; pk_prev = pk_cur; pk_cur = tp;
  }
  return GAC_OK;
}

int aiqn_sample(gac_ctx* c, const float* actor, const float* states_u, int n_states, const int* sidx, const float* taus,
                int rows, float* actions) {
  if (c->cfg.actor_kind != GAC_ACTOR_AIQN) return fail(GAC_ERR_INVALID, "gac_aiqn_sample: context was created for the IQN actor");
  if (n_states > c->NUs) return fail(GAC_ERR_WORKSPACE, "gac_aiqn_sample: n_states exceeds the workspace capacity");
  GAC_TRY(state_prep(c, actor, states_u, n_states, c->se_u, c->sx_u));
  const int A = c->cfg.A;
  {
    static const bool emb_off = getenv("GAC_NO_FUSED_EMB") != nullptr;
    c->emb_fused = !emb_off && c->cfg.engine == 0 && rows >= 64;
    if (c->emb_fused) {
      GAC_CUDA(emb_pack(var(c, actor, 0, GAC_A_WA), c->emb_sc[0], c->emb_pack_buf[0], c->stream));
      GAC_CUDA(emb_pack(var(c, actor, 0, GAC_A_WN), c->emb_sc[1], c->emb_pack_buf[1], c->stream));
      c->launches += 6;
    }
  }
  c->pk[PK_T_WA] = pack_w(c, PK_T_WA, var(c, actor, 0, GAC_A_WA), GAC_E, GAC_NB, GAC_E, 0, rows);
  c->gru_t = pack_gru(c, 0, actor, rows);
  if (!c->gru_t) {
    c->pk[PK_T_KXA] = pack_w(c, PK_T_KXA, var(c, actor, 0, GAC_A_KX) + (size_t)GAC_E * GAC_G, GAC_G, GAC_E, GAC_G, 0, rows);
    c->pk[PK_T_U] = pack_w(c, PK_T_U, var(c, actor, 0, GAC_A_U), GAC_G, GAC_E, GAC_G, 0, rows);
  }
  c->pk[PK_T_WN] = pack_w(c, PK_T_WN, var(c, actor, 0, GAC_A_WN), GAC_E, GAC_NB, GAC_E, 0, rows);
  {
    // fused head (w1 GEMM + LeakyReLU + w2 row dot in one kernel); |h * noise_embedding| is bounded by the noise embedding's weights
    static const bool off = getenv("GAC_NO_FUSED_HEAD") != nullptr;
    static const bool tf32_only = getenv("GAC_HEAD_TF32") != nullptr;
    c->head_fused = !off && c->cfg.engine == 0 && rows >= 64;
    if (c->head_fused) {
      GAC_CUDA(dot_pack(var(c, actor, 0, GAC_A_W1), GAC_E, 0, GAC_E, GAC_E, HEAD_BN, HEAD_NT, var(c, actor, 0, GAC_A_WN),
                        var(c, actor, 0, GAC_A_BN), GAC_NB, GAC_E, c->head_pack32, c->head_pack16, tf32_only ? nullptr : c->head_sc, c->stream));
      if (!tf32_only) GAC_CUDA(dot_ascale(c->head_sc, nullptr, c->stream));
      c->launches += tf32_only ? 1 : 5;
      c->pk[PK_T_W1] = nullptr;
    } else {
      c->pk[PK_T_W1] = pack_w(c, PK_T_W1, var(c, actor, 0, GAC_A_W1), GAC_E, GAC_E, GAC_E, 0, rows);
    }
  }
  for (int r0 = 0; r0 < rows; r0 += c->Rs) {
    const int n = rows - r0 < c->Rs ? rows - r0 : c->Rs;
    GAC_TRY(aiqn_sample_rows(c, actor, sidx + r0, taus + (size_t)r0 * A, n, actions + (size_t)r0 * A));
  }
  return GAC_OK;
}

int td_fwd_bwd(gac_ctx* c, const float* fnn, int n_in, const float* x, const float* y, int rows, float* grad, float* loss_out) {
  if (rows > c->cfg.B) return fail(GAC_ERR_WORKSPACE, "gac_td_fwd_bwd: rows exceeds B");
  cudaStream_t st = c->stream;
  const RowValid all{nullptr, 1};
  int64_t off[7];
  fnn_offsets(n_in, off);
  GAC_TRY(fnn_forward(c, fnn, n_in, x, rows, c->td_h0, c->td_h1));
  rowdot_kernel<<<cdiv(rows * 32, 256), 256, 0, st>>>(c->td_h1, GAC_H2, fnn + off[4], fnn + off[5], nullptr, nullptr, nullptr,
                                                     GAC_H2, rows, 0, c->td_pred, 1, nullptr, 0, all);
  LAUNCHED(c, "rowdot");
  td_dout_kernel<<<1, 1024, 0, st>>>(c->td_pred, y, rows, c->td_dout, loss_out);
  LAUNCHED(c, "td_dout");
  // layer 3
  GAC_TRY(run_colsum(c, c->td_h1, GAC_H2, c->td_dout, GAC_H2, rows, all, grad + off[4], 0));
  GAC_TRY(run_colsum(c, c->td_dout, 1, nullptr, 1, rows, all, grad + off[5], 0));
  outer_gate_kernel<<<(unsigned)cdiv64((long long)rows * GAC_H2, 256), 256, 0, st>>>(c->td_dout, fnn + off[4], c->td_h1, GAC_H2,
                                                                                  rows, 0.01f, c->td_da1, all);
  LAUNCHED(c, "outer_gate");
  // layer 2
  GemmArgs g = gemm_args(GAC_H1, GAC_H2, rows, c->td_h0, GAC_H1, c->td_da1, GAC_H2, grad + off[2], GAC_H2);
  g.transA = 1;
  GAC_TRY(run_gemm(c, g));
  GAC_TRY(run_colsum(c, c->td_da1, GAC_H2, nullptr, GAC_H2, rows, all, grad + off[3], 0));
  g = gemm_args(rows, GAC_H1, GAC_H2, c->td_da1, GAC_H2, fnn + off[2], GAC_H2, c->td_da0, GAC_H1);
  g.transB = 1; g.gate = c->td_h0; g.ldgate = GAC_H1; g.gate_alpha = 0.01f;
  GAC_TRY(run_gemm(c, g));
  // layer 1
  g = gemm_args(n_in, GAC_H1, rows, x, n_in, c->td_da0, GAC_H1, grad + off[0], GAC_H1);
  g.transA = 1;
  GAC_TRY(run_gemm(c, g));
  GAC_TRY(run_colsum(c, c->td_da0, GAC_H1, nullptr, GAC_H1, rows, all, grad + off[1], 0));
  return GAC_OK;
}

// Weight-only preparation of the actor's training pass: packed / pre-split tiles of every weight matrix its forward and
// backward GEMMs read, and the scales / bounds of the 3xFP16 variants.  Depends on the online actor alone, so one call serves
// every row chunk of a step (on c->stream).
int actor_train_prep(gac_ctx* c, const float* actor) {
  cudaStream_t st = c->stream;
  const float* kxa_w = var(c, actor, 0, GAC_A_KX) + (size_t)GAC_E * GAC_G;
  {
    // both cosine embeddings of the training forward on the one-kernel path of the sampler (emb_gemm_tc.cuh)
    static const bool emb_off = getenv("GAC_NO_FUSED_EMB") != nullptr;
    c->emb_train = !emb_off && c->cfg.engine == 0 && c->Mc >= 128;
    if (c->emb_train) {
      GAC_CUDA(emb_pack(var(c, actor, 0, GAC_A_WA), c->emb_sc[2], c->emb_pack_buf[2], st));
      GAC_CUDA(emb_pack(var(c, actor, 0, GAC_A_WN), c->emb_sc[3], c->emb_pack_buf[3], st));
      c->launches += 6;
    }
  }
  c->pk[PK_O_WA] = pack_w(c, PK_O_WA, var(c, actor, 0, GAC_A_WA), GAC_E, GAC_NB, GAC_E, 0, c->Mc);
  c->gru_o = pack_gru(c, 1, actor, c->Mc);
  if (!c->gru_o) {
    c->pk[PK_O_KXA] = pack_w(c, PK_O_KXA, kxa_w, GAC_G, GAC_E, GAC_G, 0, c->Mc);
    c->pk[PK_O_U] = pack_w(c, PK_O_U, var(c, actor, 0, GAC_A_U), GAC_G, GAC_E, GAC_G, 0, c->Mc);
  }
  c->pk[PK_O_WN] = pack_w(c, PK_O_WN, var(c, actor, 0, GAC_A_WN), GAC_E, GAC_NB, GAC_E, 0, c->Mc);
  c->pk[PK_O_W1] = pack_w(c, PK_O_W1, var(c, actor, 0, GAC_A_W1), GAC_E, GAC_E, GAC_E, 0, c->Mc);
  {
    // the three input-gradient contractions of the backward (x W^T): 3xFP16 with the A bound tracked by the producers
    static const bool off = getenv("GAC_NO_F16_DGRAD") != nullptr;
    c->dg_fused = !off && c->cfg.engine == 0 && c->Mc >= 128 && c->Mc % 128 == 0;
    // 3xFP16 for the three large weight gradients (dW1, dU, dkx_a): the X operand is pre-split into fp16 tiles once
    // (wg_pack_x_kernel) and fetched by bulk copies, so only dY goes through registers.  (3xFP16 with both operands
    // register-staged was no faster than 3xTF32: the loop is bound by the loaders' instruction work, not by the
    // tensor pipe -- tools/wgrad_timeline.py.)  GAC_NO_F16_WGRAD=1 restores 3xTF32.
    static const bool wg_off = getenv("GAC_NO_F16_WGRAD") != nullptr;
    c->wg_f16 = c->dg_fused && !wg_off;
    if (c->wg_f16) {
      GAC_CUDA(cudaMemsetAsync(c->wg_bound, 0, 4 * sizeof(unsigned int), st));
      tc::dot_amax_kernel<<<4, 256, 0, st>>>(nullptr, 0, 0, 0, var(c, actor, 0, GAC_A_WN), var(c, actor, 0, GAC_A_BN), GAC_NB, GAC_E, c->wg_bound);
      tc::dot_amax_kernel<<<4, 256, 0, st>>>(nullptr, 0, 0, 0, var(c, actor, 0, GAC_A_WA), var(c, actor, 0, GAC_A_BA), GAC_NB, GAC_E, c->wg_bound + 2);
      c->launches += 2;
    }
    if (c->dg_fused) {
      const float* W[3] = {var(c, actor, 0, GAC_A_W1), var(c, actor, 0, GAC_A_U), kxa_w};
      for (int i = 0; i < 3; ++i)
        GAC_CUDA(dot_pack(W[i], i == 0 ? GAC_E : GAC_G, 1, i == 0 ? GAC_E : GAC_G, GAC_E, HEAD_BN, HEAD_NT, nullptr, nullptr, 0, 0,
                          c->dg_pack32[i], c->dg_pack16[i], c->dg_sc[i], st));
      // forward head GEMM of the training pass: |h * noise_embedding| is bounded by the noise embedding's weights
      GAC_CUDA(dot_pack(var(c, actor, 0, GAC_A_W1), GAC_E, 0, GAC_E, GAC_E, HEAD_BN, HEAD_NT, var(c, actor, 0, GAC_A_WN), var(c, actor, 0, GAC_A_BN),
                        GAC_NB, GAC_E, c->w1f_pack32, c->w1f_pack16, c->w1f_sc, st));
      GAC_CUDA(dot_ascale(c->w1f_sc, nullptr, st));
      c->launches += 17;
      c->pk[PK_O_W1T] = c->pk[PK_O_UT] = c->pk[PK_O_KXAT] = nullptr;
    } else {
      c->pk[PK_O_W1T] = pack_w(c, PK_O_W1T, var(c, actor, 0, GAC_A_W1), GAC_E, GAC_E, GAC_E, 1, c->Mc);
      c->pk[PK_O_UT] = pack_w(c, PK_O_UT, var(c, actor, 0, GAC_A_U), GAC_G, GAC_G, GAC_E, 1, c->Mc);
      c->pk[PK_O_KXAT] = pack_w(c, PK_O_KXAT, kxa_w, GAC_G, GAC_G, GAC_E, 1, c->Mc);
    }
  }
  return GAC_OK;
}

int supervised_fwd(gac_ctx* c, const float* actor, const float* states_u, int n_states, const int* sidx, const float* taus,
                   const float* teacher, int max_rows, const int* d_rows, float* pred) {
  const int A = c->cfg.A, Mc = c->Mc;
  if (max_rows > Mc) return fail(GAC_ERR_WORKSPACE, "gac_aiqn_supervised_fwd: max_rows exceeds chunk_rows");
  if (n_states > c->NUa) return fail(GAC_ERR_WORKSPACE, "gac_aiqn_supervised_fwd: n_states exceeds the workspace capacity");
  cudaStream_t st = c->stream;
  // rows beyond max_rows inside the chunk are dead too: clamp through a device counter
  int* live = c->rows_c + c->nchunks;  // scratch slot
  if (d_rows) GAC_CUDA(cudaMemcpyAsync(live, d_rows, sizeof(int), cudaMemcpyDeviceToDevice, st));
  else GAC_CUDA(cudaMemcpyAsync(live, &max_rows, sizeof(int), cudaMemcpyHostToDevice, st));
  const RowValid rv{live, Mc};
  const long long AM = (long long)A * Mc;
  // the backward needs the row->state map and the unique states: keep private copies (the
  // caller's buffers may be gone by then)
  GAC_CUDA(cudaMemcpyAsync(c->a_sidx, sidx, (size_t)max_rows * sizeof(int), cudaMemcpyDeviceToDevice, st));
  GAC_CUDA(cudaMemcpyAsync(c->a_states, states_u, (size_t)n_states * c->cfg.S * sizeof(float), cudaMemcpyDeviceToDevice, st));
  c->bwd_sidx = c->a_sidx; c->bwd_states = c->a_states; c->bwd_nstates = n_states; c->bwd_drows = live;
  GAC_TRY(state_prep(c, actor, states_u, n_states, c->a_se_u, c->a_sx_u));
  const float* kxa_w = var(c, actor, 0, GAC_A_KX) + (size_t)GAC_E * GAC_G;
  if (c->actor_prep_token != actor) GAC_TRY(actor_train_prep(c, actor));   // else: gac_step_grads issued it once for all chunks, on its side lane
  cos_basis_seq_kernel<<<(unsigned)cdiv64(AM * (GAC_NB / 4), 256), 256, 0, st>>>(teacher, taus, A, Mc, c->ca, c->cn, live);
  LAUNCHED(c, "cos_basis_seq");
  // cosine embedding of the step-major training rows as one kernel (basis in registers); the bases ca / cn above are kept for
  // the embeddings' weight gradients only
  auto embed_train = [&](int which, const float* x, int xoff, const float* W, const float* bias, int act, float alpha, float* out, unsigned char* out16,
                         const float* mul = nullptr, float* out2 = nullptr) -> int {
    tc::EmbArgs e;
    e.rows = (int)AM; e.dyn_rows = live; e.seg_rows = Mc; e.x = x; e.incx = A; e.xseg = Mc; e.xoff = xoff; e.wpk = c->emb_pack_buf[which]; e.sc = c->emb_sc[which];
    e.W = W; e.bias = bias; e.act = act; e.alpha = alpha; e.mul = mul; e.ldmul = GAC_E; e.C = out; e.ldc = GAC_E; e.C2 = out2;
    e.C16 = out16; e.c16_scale = &c->g2_sc[1]->s_ae; e.c16_ok = &c->g2_sc[1]->ok; e.c16_also = 1;
    GAC_CUDA(launch_emb_gemm(e, st));
    c->launches++;
    return GAC_OK;
  };
  const bool ae_image_by_emb = c->emb_train && c->gru_o && c->g2_on[1];
  GemmArgs g;
  if (c->emb_train) {
    GAC_TRY(embed_train(2, teacher, -1, var(c, actor, 0, GAC_A_WA), var(c, actor, 0, GAC_A_BA), 1, 0.2f, c->a_ae, ae_image_by_emb ? c->a_ae_pk : nullptr));
  } else {
    g = gemm_args((int)AM, GAC_E, GAC_NB, c->ca, GAC_NB, var(c, actor, 0, GAC_A_WA), GAC_E, c->a_ae, GAC_E);
    g.bias = var(c, actor, 0, GAC_A_BA); g.act = ACT_LRELU; g.alpha = 0.2f; g.dyn_rows = live; g.seg_rows = Mc; g.Bpacked = c->pk[PK_O_WA]; g.Bpacked_bn = c->pk_bn[PK_O_WA];
    GAC_TRY(run_gemm(c, g));
  }
  if (!c->gru_o) {
    g = gemm_args((int)AM, GAC_G, GAC_E, c->a_ae, GAC_E, kxa_w, GAC_G, c->mx, GAC_G);
    g.dyn_rows = live; g.seg_rows = Mc; g.Bpacked = c->pk[PK_O_KXA]; g.Bpacked_bn = c->pk_bn[PK_O_KXA];
    GAC_TRY(run_gemm(c, g));
  }
  for (int d = 0; d < A; ++d) {
    const size_t o4 = (size_t)d * Mc * GAC_E, o12 = (size_t)d * Mc * GAC_G;
    if (c->gru_o) {
      const bool v2 = c->g2_on[1];
      if (v2) {
        if (d == 0 && !ae_image_by_emb) {   // every step's action embedding -> a16 images in one pass (row blocks never straddle a step: Mc is a multiple of 128)
          GAC_CUDA(a16_pack(c->a_ae, GAC_E, AM, live, Mc, &c->g2_sc[1]->s_ae, c->a_ae_pk, st));
          c->launches++;
        }
        GAC_TRY(run_gru_step_v2(c, 1, actor, c->a_sx_u, sidx, c->a_ae_pk + tc::a16_bytes((long long)d * Mc), c->a_ae + o4, c->a_h_pk[d & 1],
                                d ? c->hall + o4 : nullptr, c->hall + o4 + (size_t)Mc * GAC_E, c->a_h_pk[(d + 1) & 1], Mc, live, c->zs + o4,
                                c->rs + o4, c->cs + o4, c->mhh + o4));
      } else {
        GAC_TRY(run_gru_step(c, 1, actor, c->a_sx_u, sidx, c->a_ae + o4, d ? c->hall + o4 : nullptr, c->hall + o4 + (size_t)Mc * GAC_E, Mc,
                             live, c->zs + o4, c->rs + o4, c->cs + o4, c->mhh + o4, true, nullptr));
      }
      continue;
    }
    if (d) {
      g = gemm_args(Mc, GAC_G, GAC_E, c->hall + o4, GAC_E, var(c, actor, 0, GAC_A_U), GAC_G, c->dmh + o12, GAC_G);
      g.dyn_rows = live; g.seg_rows = Mc; g.Bpacked = c->pk[PK_O_U]; g.Bpacked_bn = c->pk_bn[PK_O_U];
      GAC_TRY(run_gemm(c, g));
    }
    gru_gates_fwd_kernel<<<(unsigned)cdiv64((long long)Mc * GAC_E, 256), 256, 0, st>>>(
        c->a_sx_u, sidx, c->mx + o12, d ? c->dmh + o12 : nullptr, var(c, actor, 0, GAC_A_GB) + GAC_G, c->hall + o4,
        c->hall + o4 + (size_t)Mc * GAC_E, c->zs + o4, c->rs + o4, c->cs + o4, c->mhh + o4, Mc, rv);
    LAUNCHED(c, "gru_gates_fwd");
  }
  if (c->emb_train) {
    // noise embedding and its Hadamard product with the GRU states of every step in one pass (ne is kept for the backward)
    GAC_TRY(embed_train(3, taus, 0, var(c, actor, 0, GAC_A_WN), var(c, actor, 0, GAC_A_BN), 0, 0.f, c->ne, nullptr, c->hall + (size_t)Mc * GAC_E, c->au));
  } else {
    g = gemm_args((int)AM, GAC_E, GAC_NB, c->cn, GAC_NB, var(c, actor, 0, GAC_A_WN), GAC_E, c->ne, GAC_E);
    g.bias = var(c, actor, 0, GAC_A_BN); g.dyn_rows = live; g.seg_rows = Mc; g.Bpacked = c->pk[PK_O_WN]; g.Bpacked_bn = c->pk_bn[PK_O_WN];
    GAC_TRY(run_gemm(c, g));
  }
  if (!c->emb_train) {
    mul_kernel<<<(unsigned)cdiv64(AM * GAC_E / 4, 256), 256, 0, st>>>(c->ne, c->hall + (size_t)Mc * GAC_E, AM * GAC_E, Mc, live, c->au);
    LAUNCHED(c, "hadamard");
  }
  if (c->dg_fused) {
    tc::DotGemmArgs q;
    q.rows = (int)AM; q.dyn_rows = live; q.seg_rows = Mc; q.A = c->au; q.lda = GAC_E; q.K = GAC_E; q.N = GAC_E; q.bn = HEAD_BN; q.ntiles = HEAD_NT;
    q.wpk = c->w1f_pack32; q.wpk16 = c->w1f_pack16; q.sc = c->w1f_sc;
    q.bias = var(c, actor, 0, GAC_A_B1); q.w2 = nullptr; q.alpha = 0.01f; q.part = nullptr;
    q.mode = 1; q.C = c->ay1; q.ldc = GAC_E; q.gate = nullptr; q.ldgate = 0; q.gate_alpha = 0.f; q.accumulate = 0; q.act_lrelu = 1; q.dbg = nullptr;
    GAC_CUDA(launch_dot_gemm(q, st));
    c->launches++;
  } else {
    g = gemm_args((int)AM, GAC_E, GAC_E, c->au, GAC_E, var(c, actor, 0, GAC_A_W1), GAC_E, c->ay1, GAC_E);
    g.bias = var(c, actor, 0, GAC_A_B1); g.act = ACT_LRELU; g.alpha = 0.01f; g.dyn_rows = live; g.seg_rows = Mc; g.Bpacked = c->pk[PK_O_W1]; g.Bpacked_bn = c->pk_bn[PK_O_W1];
    GAC_TRY(run_gemm(c, g));
  }
  rowdot_steps_kernel<<<dim3(cdiv(Mc * 32, 256), A), 256, 0, st>>>(c->ay1, GAC_E, var(c, actor, 0, GAC_A_W2), var(c, actor, 0, GAC_A_B2), GAC_E, Mc, A,
                                                                  c->pred_c, pred, rv);
  LAUNCHED(c, "rowdot_tanh");
  return GAC_OK;
}

// Lanes of gac_step_grads (lanes_on): work that nothing on the critical chain waits for goes to the side stream with its own
// split-K / column-sum scratch.  Outside gac_step_grads (the single-op C-ABI entries) these are no-ops.
inline void lane_side(gac_ctx* c) { if (c->lanes_on) { c->stream = c->side; c->spart = c->spart2; c->cpart = c->cpart2; } }
inline void lane_main(gac_ctx* c) { if (c->lanes_on) { c->stream = c->main_stream; c->spart = c->spart_main; c->cpart = c->cpart_main; } }
// side lane continues only after everything issued on the main lane so far
inline cudaError_t lane_side_after_main(gac_ctx* c) {
  if (!c->lanes_on) return cudaSuccess;
  cudaError_t e = cudaEventRecord(c->ev_w, c->main_stream);
  return e != cudaSuccess ? e : cudaStreamWaitEvent(c->side, c->ev_w, 0);
}
inline cudaError_t lane_main_after_side(gac_ctx* c) {
  if (!c->lanes_on) return cudaSuccess;
  cudaError_t e = cudaEventRecord(c->ev_wjoin, c->side);
  return e != cudaSuccess ? e : cudaStreamWaitEvent(c->main_stream, c->ev_wjoin, 0);
}

int aiqn_bwd(gac_ctx* c, const float* actor, const float* dpred, float* grad, int acc) {
  const int A = c->cfg.A, Mc = c->Mc, S = c->cfg.S;
  if (!c->bwd_drows) return fail(GAC_ERR_INVALID, "gac_aiqn_bwd: no saved forward");
  cudaStream_t st = c->stream;
  const int* live = c->bwd_drows;
  const RowValid rv{live, Mc};
  const long long AM = (long long)A * Mc;
  const float* kxa = var(c, actor, 0, GAC_A_KX) + (size_t)GAC_E * GAC_G;
  auto G = [&](int v) { return var(c, grad, 0, v); };
  // site >= 0: 3xFP16 with |dY| from the producers' running maximum (amax_dy) and |X| <= *bound_x (or bound_const)
  auto wgrad = [&](int M, int N, const float* X, int ldx, const float* dY, int ldy, long long rows, float* dst, bool dyn, int site = -1,
                   const unsigned int* amax_dy = nullptr, const unsigned int* bound_x = nullptr, float bound_const = 1.f) -> int {
    GemmArgs g = gemm_args(M, N, (int)rows, X, ldx, dY, ldy, dst, N);
    g.transA = 1; g.accumulate = acc;
    if (site >= 0 && c->wg_f16) {
      wg_scale_kernel<<<1, 1, 0, c->stream>>>(c->wg_sc + site, amax_dy, bound_x, bound_const);
      LAUNCHED(c, "wg_scale");
      g.f16 = c->wg_sc + site;
      if (dyn && M == GAC_E && ldx == GAC_E && rows == AM && c->cfg.engine == 0) {
        GAC_CUDA(wg_pack_x(X, ldx, M, Mc, A, live, c->wg_sc + site, c->wg_xpack, c->stream));
        c->launches++;
        g.Bpk16 = c->wg_xpack;
      }
    }
    static const bool timeline = getenv("GAC_WGRAD_TIMELINE") != nullptr;
    if (timeline && site == 2) g.dbg = c->dbg_buf;           // the kx_a weight gradient: (2, 10, splits) CTAs
    if (dyn) { g.dyn_rows = live; g.seg_rows = Mc; g.valid_on_k = 1; }
    return run_gemm(c, g);
  };
  // The five time-parallel weight gradients feed nothing downstream: each is issued on the side lane as soon as its operands
  // exist (an event orders it after the main lane's producers) and runs beside the serial chain head -> Hadamard ->
  // recurrence -> input gradients, much of which is HBM-bound.  Joined at the end of this function: the next chunk's forward
  // overwrites their operands.
  auto side_wgrad = [&](auto&& fn) -> int {
    GAC_CUDA(lane_side_after_main(c));
    lane_side(c);
    const int rc = fn();
    lane_main(c);
    return rc;
  };
  // C (rows, 400) = X (rows, K) * W^T through dot_gemm_tc.cuh (store mode); which: 0 = W1, 1 = U, 2 = kx_a
  auto dgrad = [&](int which, const float* X, int ldx, int K, long long rows, float* Cdst, const float* gate, float gate_alpha, int accumulate,
                   const unsigned int* amax) -> int {
    GAC_CUDA(dot_ascale(c->dg_sc[which], amax, st));
    tc::DotGemmArgs q;
    q.rows = (int)rows; q.dyn_rows = live; q.seg_rows = Mc; q.A = X; q.lda = ldx; q.K = K; q.N = GAC_E; q.bn = HEAD_BN; q.ntiles = HEAD_NT;
    static const bool dg_tf32 = getenv("GAC_DGRAD_TF32") != nullptr;   // A/B switch: same pipeline, 3xTF32 arithmetic
    q.wpk = c->dg_pack32[which]; q.wpk16 = c->dg_pack16[which]; q.sc = dg_tf32 ? nullptr : c->dg_sc[which];
    q.bias = nullptr; q.w2 = nullptr; q.alpha = 0.f; q.part = nullptr;
    q.mode = 1; q.C = Cdst; q.ldc = GAC_E; q.gate = gate; q.ldgate = GAC_E; q.gate_alpha = gate_alpha; q.accumulate = accumulate; q.act_lrelu = 0;
    static const bool dg_timeline = getenv("GAC_DGRAD_TIMELINE") != nullptr;
    q.dbg = (dg_timeline && which == 2) ? c->dbg_buf : nullptr;   // the kx_a input gradient
    GAC_CUDA(launch_dot_gemm(q, st));
    c->launches += 2;
    return GAC_OK;
  };
  if (c->dg_fused) GAC_CUDA(cudaMemsetAsync(c->dg_amax, 0, 4 * sizeof(unsigned int), st));
  unsigned int* amax_g = c->dg_fused ? c->dg_amax : nullptr;
  // head
  head_dop_kernel<<<(unsigned)cdiv64(AM, 256), 256, 0, st>>>(dpred, c->pred_c, A, Mc, live, c->dop);
  LAUNCHED(c, "head_dop");
  GAC_TRY(run_colsum(c, c->dop, 1, nullptr, 1, AM, rv, G(GAC_A_B2), acc));
  const int nrb_all = (int)cdiv64(AM, FB_ROWS);
  head_bwd_kernel<<<dim3(GAC_E / FB_COLS, nrb_all), FB_COLS, 0, st>>>(c->dop, var(c, actor, 0, GAC_A_W2), c->ay1, AM, rv, c->dy1p,
                                                                      c->fpart0, c->fpart1, amax_g ? amax_g + 2 : nullptr);
  LAUNCHED(c, "head_bwd");
  GAC_TRY(run_reduce_cols(c, c->fpart0, nrb_all, GAC_E, G(GAC_A_B1), acc));
  GAC_TRY(run_reduce_cols(c, c->fpart1, nrb_all, GAC_E, G(GAC_A_W2), acc));
  GAC_TRY(side_wgrad([&] { return wgrad(GAC_E, GAC_E, c->au, GAC_E, c->dy1p, GAC_E, AM, G(GAC_A_W1), true, 0, c->dg_amax + 2, c->wg_bound + 1); }));
  GemmArgs g = gemm_args((int)AM, GAC_E, GAC_E, c->dy1p, GAC_E, var(c, actor, 0, GAC_A_W1), GAC_E, c->du, GAC_E);
  if (c->dg_fused) {
    GAC_TRY(dgrad(0, c->dy1p, GAC_E, GAC_E, AM, c->du, nullptr, 0.f, 0, amax_g + 2));
  } else {
    g.transB = 1; g.dyn_rows = live; g.seg_rows = Mc; g.Bpacked = c->pk[PK_O_W1T]; g.Bpacked_bn = c->pk_bn[PK_O_W1T];
    GAC_TRY(run_gemm(c, g));
  }
  hadamard_bwd_kernel<<<dim3(GAC_E / FB_COLS, nrb_all), FB_COLS, 0, st>>>(c->du, c->hall + (size_t)Mc * GAC_E, c->ne, AM, rv, c->dne,
                                                                          c->dhu, c->fpart0);
  LAUNCHED(c, "hadamard_bwd");
  GAC_TRY(run_reduce_cols(c, c->fpart0, nrb_all, GAC_E, G(GAC_A_BN), acc));
  GAC_TRY(side_wgrad([&] { return wgrad(GAC_NB, GAC_E, c->cn, GAC_NB, c->dne, GAC_E, AM, G(GAC_A_WN), true); }));
  // recurrence, reverse time.  mx is overwritten with dmx step by step (it is dead after the forward gates)
  float* dh_next = nullptr; float* carry = c->dhc0;
  const int nrb_step = Mc / FB_ROWS;
  for (int d = A - 1; d >= 0; --d) {
    const size_t o4 = (size_t)d * Mc * GAC_E, o12 = (size_t)d * Mc * GAC_G;
    gru_gates_bwd_kernel<<<dim3(GAC_E / FB_COLS, nrb_step), FB_COLS, 0, st>>>(
        dh_next, c->dhu + o4, c->zs + o4, c->rs + o4, c->cs + o4, c->mhh + o4, d ? c->hall + o4 : nullptr, c->mx + o12, c->dmh + o12,
        carry, Mc, rv, c->fpart0 + (size_t)d * nrb_step * GAC_G, c->fpart1 + (size_t)d * nrb_step * GAC_G, amax_g);
    LAUNCHED(c, "gru_gates_bwd");
    if (d && c->dg_fused) {
      // running maximum over the steps done so far: an upper bound of this step's |dmh|
      GAC_TRY(dgrad(1, c->dmh + o12, GAC_G, GAC_G, Mc, carry, nullptr, 0.f, 1, amax_g));
    } else if (d) {
      g = gemm_args(Mc, GAC_E, GAC_G, c->dmh + o12, GAC_G, var(c, actor, 0, GAC_A_U), GAC_G, carry, GAC_E);
      g.transB = 1; g.accumulate = 1; g.dyn_rows = live; g.seg_rows = Mc; g.Bpacked = c->pk[PK_O_UT]; g.Bpacked_bn = c->pk_bn[PK_O_UT];
      GAC_TRY(run_gemm(c, g));
    }
    dh_next = carry;
    carry = carry == c->dhc0 ? c->dhc1 : c->dhc0;
  }
  // time-parallel weight gradients of the GRU
  GAC_TRY(side_wgrad([&] { return wgrad(GAC_E, GAC_G, c->hall, GAC_E, c->dmh, GAC_G, AM, G(GAC_A_U), true, 1, c->dg_amax, nullptr, 1.f); }));
  GAC_TRY(run_reduce_cols(c, c->fpart1, A * nrb_step, GAC_G, G(GAC_A_GB) + GAC_G, acc));
  GAC_TRY(run_reduce_cols(c, c->fpart0, A * nrb_step, GAC_G, G(GAC_A_GB), acc));
  GAC_TRY(side_wgrad([&] { return wgrad(GAC_E, GAC_G, c->a_ae, GAC_E, c->mx, GAC_G, AM, G(GAC_A_KX) + (size_t)GAC_E * GAC_G, true, 2, c->dg_amax, c->wg_bound + 3); }));
  if (c->dg_fused) {
    GAC_TRY(dgrad(2, c->mx, GAC_G, GAC_G, AM, c->du, c->a_ae, 0.2f, 0, amax_g));
  } else {
    g = gemm_args((int)AM, GAC_E, GAC_G, c->mx, GAC_G, kxa, GAC_G, c->du, GAC_E);
    g.transB = 1; g.gate = c->a_ae; g.ldgate = GAC_E; g.gate_alpha = 0.2f; g.dyn_rows = live; g.seg_rows = Mc; g.Bpacked = c->pk[PK_O_KXAT]; g.Bpacked_bn = c->pk_bn[PK_O_KXAT];
    GAC_TRY(run_gemm(c, g));
  }
  GAC_TRY(side_wgrad([&] {
    GAC_TRY(wgrad(GAC_NB, GAC_E, c->ca, GAC_NB, c->du, GAC_E, AM, G(GAC_A_WA), true));
    return run_colsum(c, c->du, GAC_E, nullptr, GAC_E, AM, rv, G(GAC_A_BA), acc);
  }));
  // state half: reduce d(sx) per unique state, then B-row GEMMs
  const int NU = c->bwd_nstates;
  // rows grouped by state with a stable counting sort (block histograms -> scan -> ranked scatter), then one block
  // per (state, column slab) streams exactly its rows in ascending row order: deterministic, no atomics on floats
  const int nsb = cdiv(Mc, SEG_BLOCK);
  {
    // one int of shared memory per unique state: above the 48 KB default the opt-in limit has to be raised (derive() caps B)
    static size_t hist_attr = 48 * 1024;
    if (NU * sizeof(int) > hist_attr) {
      GAC_CUDA(cudaFuncSetAttribute(seg_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NU * sizeof(int))));
      hist_attr = NU * sizeof(int);
    }
  }
  seg_hist_kernel<<<nsb, SEG_BLOCK, NU * sizeof(int), st>>>(c->bwd_sidx, Mc, live, NU, c->seg_hist);
  LAUNCHED(c, "seg_hist");
  seg_scan_kernel<<<1, 1024, 0, st>>>(c->seg_hist, nsb, NU, c->seg_info);
  LAUNCHED(c, "seg_scan");
  seg_scatter_kernel<<<nsb, SEG_BLOCK, 0, st>>>(c->bwd_sidx, Mc, live, NU, c->seg_hist, c->seg_perm);
  LAUNCHED(c, "seg_scatter");
  segment_sum_kernel<<<dim3(NU, cdiv(GAC_G, 256)), 256, 0, st>>>(c->mx, c->seg_perm, c->seg_info, A, Mc, NU, 0, c->a_dsx_u);
  LAUNCHED(c, "segment_sum");
  GAC_TRY(wgrad(GAC_E, GAC_G, c->a_se_u, GAC_E, c->a_dsx_u, GAC_G, NU, G(GAC_A_KX), false));
  g = gemm_args(NU, GAC_E, GAC_G, c->a_dsx_u, GAC_G, var(c, actor, 0, GAC_A_KX), GAC_G, c->a_dse_u, GAC_E);
  g.transB = 1; g.gate = c->a_se_u; g.ldgate = GAC_E; g.gate_alpha = 0.01f * 0.2f;
  GAC_TRY(run_gemm(c, g));
  GAC_TRY(wgrad(S, GAC_E, c->bwd_states, S, c->a_dse_u, GAC_E, NU, G(GAC_A_WS), false));
  GAC_TRY(run_colsum(c, c->a_dse_u, GAC_E, nullptr, GAC_E, NU, RowValid{nullptr, 1}, G(GAC_A_BS), acc));
  GAC_CUDA(lane_main_after_side(c));
  return GAC_OK;
}

// ---------------------------------- IQN actor (StochasticActor) ----------------------------------
inline const float* ivar(const gac_ctx* c, const float* base, int v) { return base + c->lay.offsets[v]; }
inline float* ivar(const gac_ctx* c, float* base, int v) { return base + c->lay.offsets[v]; }

int iqn_forward(gac_ctx* c, const float* actor, const float* states_u, int n_states, const int* sidx, const float* taus, int rows,
                float* actions) {
  if (c->cfg.actor_kind != GAC_ACTOR_IQN) return fail(GAC_ERR_INVALID, "gac_iqn_forward: context was not created for the IQN actor");
  if (n_states > c->NUs) return fail(GAC_ERR_WORKSPACE, "gac_iqn_forward: n_states exceeds the workspace capacity");
  const int S = c->cfg.S, A = c->cfg.A, d = c->iq_d, D = c->iq_D;
  cudaStream_t st = c->stream;
  const RowValid all{nullptr, 1};
  // state embedding once per unique state (exact hoist), gathered inside the Hadamard kernel
  GemmArgs g = gemm_args(n_states, D, S, states_u, S, ivar(c, actor, GAC_I_WS), D, c->iq_se_u, D);
  g.bias = ivar(c, actor, GAC_I_BS); g.act = ACT_LRELU; g.alpha = 0.01f;
  GAC_TRY(run_gemm(c, g));
  for (int r0 = 0; r0 < rows; r0 += c->Rs) {
    const int n = rows - r0 < c->Rs ? rows - r0 : c->Rs;
    cos_basis_kernel<<<(unsigned)cdiv64((long long)n * A * (GAC_NB / 4), 256), 256, 0, st>>>(taus + (size_t)r0 * A, 1, c->iq_cos, n * A, all);
    LAUNCHED(c, "cos_basis");
    g = gemm_args(n * A, d, GAC_NB, c->iq_cos, GAC_NB, ivar(c, actor, GAC_I_WN), d, c->iq_ne, d);
    g.bias = ivar(c, actor, GAC_I_BN); g.act = ACT_LRELU; g.alpha = 0.01f;
    GAC_TRY(run_gemm(c, g));
    iqn_merge_kernel<<<(unsigned)cdiv64((long long)n * D, 256), 256, 0, st>>>(c->iq_se_u, sidx + r0, c->iq_ne, n, D, all, c->iq_ne);
    LAUNCHED(c, "iqn_merge");
    g = gemm_args(n, 200, D, c->iq_ne, D, ivar(c, actor, GAC_I_WM), 200, c->iq_me, 200);
    g.bias = ivar(c, actor, GAC_I_BM); g.act = ACT_LRELU; g.alpha = 0.01f;
    GAC_TRY(run_gemm(c, g));
    g = gemm_args(n, A, 200, c->iq_me, 200, ivar(c, actor, GAC_I_WO), A, actions + (size_t)r0 * A, A);
    g.bias = ivar(c, actor, GAC_I_BO); g.act = ACT_TANH;
    GAC_TRY(run_gemm(c, g));
  }
  return GAC_OK;
}

int iqn_train_fwd(gac_ctx* c, const float* actor, const float* states_u, int n_states, const int* sidx, const float* taus, int max_rows,
                  const int* d_rows, float* pred) {
  if (c->cfg.actor_kind != GAC_ACTOR_IQN) return fail(GAC_ERR_INVALID, "gac_iqn_train_fwd: context was not created for the IQN actor");
  const int S = c->cfg.S, A = c->cfg.A, d = c->iq_d, D = c->iq_D, Mc = c->Mc;
  if (max_rows > Mc) return fail(GAC_ERR_WORKSPACE, "gac_iqn_train_fwd: max_rows exceeds chunk_rows");
  cudaStream_t st = c->stream;
  int* live = c->it_live;
  if (d_rows) GAC_CUDA(cudaMemcpyAsync(live + 2, d_rows, sizeof(int), cudaMemcpyDeviceToDevice, st));
  else GAC_CUDA(cudaMemcpyAsync(live + 2, &max_rows, sizeof(int), cudaMemcpyHostToDevice, st));
  scale_int_kernel<<<1, 32, 0, st>>>(live + 2, A, max_rows, live);
  LAUNCHED(c, "scale_int");
  const RowValid rv{live, Mc}, rvA{live + 1, Mc * A};
  gather_rows_kernel<<<(unsigned)cdiv64((long long)max_rows * S, 256), 256, 0, st>>>(states_u, sidx, max_rows, S, rv, c->it_x);
  LAUNCHED(c, "gather_rows");
  GemmArgs g = gemm_args(Mc, D, S, c->it_x, S, ivar(c, actor, GAC_I_WS), D, c->it_se, D);
  g.bias = ivar(c, actor, GAC_I_BS); g.act = ACT_LRELU; g.alpha = 0.01f; g.dyn_rows = live; g.seg_rows = Mc;
  GAC_TRY(run_gemm(c, g));
  cos_basis_kernel<<<(unsigned)cdiv64((long long)max_rows * A * (GAC_NB / 4), 256), 256, 0, st>>>(taus, 1, c->it_cos, max_rows * A, rvA);
  LAUNCHED(c, "cos_basis");
  g = gemm_args(Mc * A, d, GAC_NB, c->it_cos, GAC_NB, ivar(c, actor, GAC_I_WN), d, c->it_ne, d);
  g.bias = ivar(c, actor, GAC_I_BN); g.act = ACT_LRELU; g.alpha = 0.01f; g.dyn_rows = live + 1; g.seg_rows = Mc * A;
  GAC_TRY(run_gemm(c, g));
  iqn_merge_kernel<<<(unsigned)cdiv64((long long)Mc * D, 256), 256, 0, st>>>(c->it_se, nullptr, c->it_ne, Mc, D, rv, c->it_merge);
  LAUNCHED(c, "iqn_merge");
  g = gemm_args(Mc, 200, D, c->it_merge, D, ivar(c, actor, GAC_I_WM), 200, c->it_me, 200);
  g.bias = ivar(c, actor, GAC_I_BM); g.act = ACT_LRELU; g.alpha = 0.01f; g.dyn_rows = live; g.seg_rows = Mc;
  GAC_TRY(run_gemm(c, g));
  g = gemm_args(Mc, A, 200, c->it_me, 200, ivar(c, actor, GAC_I_WO), A, c->pred_c, A);
  g.bias = ivar(c, actor, GAC_I_BO); g.act = ACT_TANH; g.dyn_rows = live; g.seg_rows = Mc;
  GAC_TRY(run_gemm(c, g));
  if (pred) GAC_CUDA(cudaMemcpyAsync(pred, c->pred_c, (size_t)max_rows * A * sizeof(float), cudaMemcpyDeviceToDevice, st));
  c->bwd_drows = live;
  return GAC_OK;
}

int iqn_bwd(gac_ctx* c, const float* actor, const float* dpred, float* grad, int acc) {
  if (c->cfg.actor_kind != GAC_ACTOR_IQN || !c->bwd_drows) return fail(GAC_ERR_INVALID, "gac_iqn_bwd: no saved IQN forward");
  const int S = c->cfg.S, A = c->cfg.A, d = c->iq_d, D = c->iq_D, Mc = c->Mc;
  cudaStream_t st = c->stream;
  const int* live = c->it_live;
  const RowValid rv{live, Mc}, rvA{live + 1, Mc * A};
  auto wgrad = [&](int M, int N, const float* X, int ldx, const float* dY, int ldy, int rows, float* dst, const int* dyn, int seg) {
    GemmArgs g = gemm_args(M, N, rows, X, ldx, dY, ldy, dst, N);
    g.transA = 1; g.accumulate = acc; g.dyn_rows = dyn; g.seg_rows = seg; g.valid_on_k = 1;
    return run_gemm(c, g);
  };
  tanh_bwd_kernel<<<(unsigned)cdiv64((long long)Mc * A, 256), 256, 0, st>>>(dpred, c->pred_c, (long long)Mc * A, A, rv, c->it_dop);
  LAUNCHED(c, "tanh_bwd");
  GAC_TRY(wgrad(200, A, c->it_me, 200, c->it_dop, A, Mc, ivar(c, grad, GAC_I_WO), live, Mc));
  GAC_TRY(run_colsum(c, c->it_dop, A, nullptr, A, Mc, rv, ivar(c, grad, GAC_I_BO), acc));
  GemmArgs g = gemm_args(Mc, 200, A, c->it_dop, A, ivar(c, actor, GAC_I_WO), A, c->it_dme, 200);
  g.transB = 1; g.gate = c->it_me; g.ldgate = 200; g.gate_alpha = 0.01f; g.dyn_rows = live; g.seg_rows = Mc;
  GAC_TRY(run_gemm(c, g));
  GAC_TRY(wgrad(D, 200, c->it_merge, D, c->it_dme, 200, Mc, ivar(c, grad, GAC_I_WM), live, Mc));
  GAC_TRY(run_colsum(c, c->it_dme, 200, nullptr, 200, Mc, rv, ivar(c, grad, GAC_I_BM), acc));
  g = gemm_args(Mc, D, 200, c->it_dme, 200, ivar(c, actor, GAC_I_WM), 200, c->it_dmerge, D);
  g.transB = 1; g.dyn_rows = live; g.seg_rows = Mc;
  GAC_TRY(run_gemm(c, g));
  iqn_dmerge_kernel<<<(unsigned)cdiv64((long long)Mc * D, 256), 256, 0, st>>>(c->it_dmerge, c->it_se, c->it_ne, (long long)Mc * D, D, rv,
                                                                            c->it_dne, c->it_dse);
  LAUNCHED(c, "iqn_dmerge");
  GAC_TRY(wgrad(GAC_NB, d, c->it_cos, GAC_NB, c->it_dne, d, Mc * A, ivar(c, grad, GAC_I_WN), live + 1, Mc * A));
  GAC_TRY(run_colsum(c, c->it_dne, d, nullptr, d, (long long)Mc * A, rvA, ivar(c, grad, GAC_I_BN), acc));
  GAC_TRY(wgrad(S, D, c->it_x, S, c->it_dse, D, Mc, ivar(c, grad, GAC_I_WS), live, Mc));
  GAC_TRY(run_colsum(c, c->it_dse, D, nullptr, D, Mc, rv, ivar(c, grad, GAC_I_BS), acc));
  return GAC_OK;
}

int compact(gac_ctx* c, const float* q, const float* v, int n, int64_t* idx_out, int* pos_out, int* m_out, float* sum_adv_out,
            float* adv_out, const float* actions, int A, float* actions_out, const int* sidx, int* sidx_out) {
  if (n > 2 * c->P) return fail(GAC_ERR_WORKSPACE, "gac_advantage_compact: n exceeds 2*B*K");
  cudaStream_t st = c->stream;
  const int nblk = cdiv(n, CMP_BLOCK);
  compact_count_kernel<<<nblk, CMP_BLOCK, 0, st>>>(q, v, n, c->blk_count);
  LAUNCHED(c, "compact_count");
  compact_scan_kernel<<<1, 1024, 0, st>>>(c->blk_count, nblk, m_out);
  LAUNCHED(c, "compact_scan");
  compact_scatter_kernel<<<nblk, CMP_BLOCK, 0, st>>>(q, v, n, c->blk_count, idx_out, pos_out, adv_out, actions, A, actions_out, sidx,
                                                     sidx_out, c->blk_adv);
  LAUNCHED(c, "compact_scatter");
  final_sum_kernel<<<1, 1024, 0, st>>>(c->blk_adv, nblk, sum_adv_out, 0);
  LAUNCHED(c, "final_sum");
  return GAC_OK;
}

int qhuber(gac_ctx* c, const float* pred, const float* target, const float* taus, const float* w, int max_rows, const int* d_rows,
           int A, float inv_norm, float* dpred, float* loss_out, int accumulate) {
  const long long n = (long long)max_rows * A;
  const int nblk = (int)cdiv64(n, 256);
  if (nblk > (2 * c->P * A + 255) / 256 + 1) return fail(GAC_ERR_WORKSPACE, "gac_qhuber_loss_bwd: rows exceed 2*B*K");
  qhuber_kernel<<<nblk, 256, 0, c->stream>>>(pred, target, taus, w, A, max_rows, d_rows, inv_norm, dpred, c->loss_part);
  LAUNCHED(c, "qhuber");
  final_sum_kernel<<<1, 1024, 0, c->stream>>>(c->loss_part, nblk, loss_out, accumulate);
  LAUNCHED(c, "final_sum");
  return GAC_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

const char* gac_last_error(void) { return g_err; }
int gac_abi_version(void) { return GAC_B200_ABI_VERSION; }

size_t gac_workspace_bytes(const gac_config_t* cfg) {
  if (!cfg) return 0;
  gac_ctx tmp;
  memset(&tmp, 0, sizeof(tmp));
  tmp.cfg = *cfg;
  if (derive(&tmp) != 0) return 0;
  return partition(&tmp, nullptr);
}

int gac_create(const gac_config_t* cfg, void* workspace, size_t workspace_bytes, void* stream, gac_ctx_t** out) {
  if (!cfg || !workspace || !out) return fail(GAC_ERR_INVALID, "gac_create: null argument");
  gac_ctx* c = new (std::nothrow) gac_ctx;
  if (!c) return fail(GAC_ERR_INVALID, "gac_create: out of host memory");
  memset(c, 0, sizeof(*c));
  c->cfg = *cfg;
  c->stream = reinterpret_cast<cudaStream_t>(stream);
  int rc = derive(c);
  if (rc) { delete c; return rc; }
  const size_t need = partition(c, nullptr);
  if (workspace_bytes < need) { delete c; return fail(GAC_ERR_WORKSPACE, "gac_create: workspace too small"); }
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) { delete c; return fail(GAC_ERR_INVALID, "gac_create: workspace must be 256-byte aligned"); }
  partition(c, workspace);
  // constants
  float ipi[GAC_NB];
  const float pif = 3.14159274101257324f;  // fl32(pi)
  for (int i = 0; i < GAC_NB; ++i) ipi[i] = (float)(i + 1) * pif;
  cudaError_t e = cudaMemcpyToSymbolAsync(c_ipi, ipi, sizeof(ipi), 0, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(c->hall, 0, (size_t)c->Mc * GAC_E * sizeof(float), c->stream);  // h_{-1} = 0
  if (e == cudaSuccess) e = tc_gemm_init();
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_q, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_cprep, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_aprep, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_w, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_wjoin, cudaEventDisableTiming);
  if (e != cudaSuccess) { gac_destroy(c); return fail(GAC_ERR_CUDA, "gac_create: %s", cudaGetErrorString(e)); }
  *out = c;
  return GAC_OK;
}

void gac_destroy(gac_ctx_t* ctx) {
  if (!ctx) return;
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_q) cudaEventDestroy(ctx->ev_q);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->ev_cprep) cudaEventDestroy(ctx->ev_cprep);
  if (ctx->ev_aprep) cudaEventDestroy(ctx->ev_aprep);
  if (ctx->ev_w) cudaEventDestroy(ctx->ev_w);
  if (ctx->ev_wjoin) cudaEventDestroy(ctx->ev_wjoin);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  delete ctx;
}
int gac_set_stream(gac_ctx_t* ctx, void* stream) {
  if (!ctx) return fail(GAC_ERR_INVALID, "gac_set_stream: null context");
  ctx->stream = reinterpret_cast<cudaStream_t>(stream);
  return GAC_OK;
}
int64_t gac_launch_count(const gac_ctx_t* ctx) { return ctx ? ctx->launches : 0; }

int gac_polyak(float* target, const float* source, int64_t n, float rho, void* stream) {
  if (!target || !source || n < 0) return fail(GAC_ERR_INVALID, "gac_polyak: bad arguments");
  if (n == 0) return GAC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float omr = (float)(1.0 - (double)rho);  // helpers.py:86: (1.0 - rate) in Python float, then fp32
  const bool aligned = ((reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(source)) & 15) == 0;
  if (aligned) polyak_kernel<<<(unsigned)cdiv64(cdiv64(n, 4), 256), 256, 0, st>>>(target, source, n, omr, rho);
  else polyak_scalar_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(target, source, n, omr, rho);
  GAC_LAUNCH_CHECK("polyak");
  return GAC_OK;
}

int gac_adam_polyak(gac_ctx_t* c, float* online, float* target, float* adam_m, float* adam_v, const float* grad,
                    const gac_hparams_t* hp, int32_t* adam_t, const float* grad_scale, const int32_t* skip) {
  if (!c || !online || !target || !adam_m || !adam_v || !grad || !hp || !adam_t) return fail(GAC_ERR_INVALID, "gac_adam_polyak: null argument");
  cudaStream_t st = c->stream;
  adam_prep_kernel<<<1, 32, 0, st>>>(adam_t, skip, hp->lr, hp->beta1, hp->beta2, c->alpha);
  LAUNCHED(c, "adam_prep");
  AdamNets nets;
  for (int i = 0; i <= GAC_NNETS; ++i) nets.begin[i] = c->lay.net_begin[i];
  const int64_t n4 = c->lay.net_begin[GAC_NNETS] / 4;
  const float omr = (float)(1.0 - (double)hp->rho);
  adam_polyak_kernel<<<(unsigned)cdiv64(n4, 256), 256, 0, st>>>(
      reinterpret_cast<float4*>(online), reinterpret_cast<float4*>(target), reinterpret_cast<float4*>(adam_m),
      reinterpret_cast<float4*>(adam_v), reinterpret_cast<const float4*>(grad), nets, c->alpha, grad_scale, skip, 1.f - hp->beta1,
      1.f - hp->beta2, hp->eps, omr, hp->rho);
  LAUNCHED(c, "adam_polyak");
  return GAC_OK;
}

int gac_aiqn_sample(gac_ctx_t* c, const float* actor, const float* states_u, int32_t n_states, const int32_t* state_idx,
                    const float* taus, int32_t rows, float* actions) {
  if (!c || !actor || !states_u || !state_idx || !taus || !actions || rows < 0) return fail(GAC_ERR_INVALID, "gac_aiqn_sample: bad arguments");
  return aiqn_sample(c, actor, states_u, n_states, state_idx, taus, rows, actions);
}

int gac_gru_step(gac_ctx_t* c, const float* actor, const float* states_u, int32_t n_states, const int32_t* state_idx,
                 const float* action_emb, const float* h_prev, int32_t rows, float* h_out, int32_t flags) {
  const bool reuse = flags & GAC_GRU_REUSE;
  if (!c || !actor || !states_u || !state_idx || !action_emb || !h_out || rows < 0) return fail(GAC_ERR_INVALID, "gac_gru_step: bad arguments");
  if (c->cfg.actor_kind != GAC_ACTOR_AIQN) return fail(GAC_ERR_INVALID, "gac_gru_step: context was created for the IQN actor");
  if (n_states > c->NUs) return fail(GAC_ERR_WORKSPACE, "gac_gru_step: n_states exceeds the workspace capacity");
  if (rows > c->Rs) return fail(GAC_ERR_WORKSPACE, "gac_gru_step: rows exceeds the sampler chunk");
  if (rows == 0) return GAC_OK;
  cudaStream_t st = c->stream;
  const float* kxa = var(c, actor, 0, GAC_A_KX) + (size_t)GAC_E * GAC_G;
  if (!reuse) {
    GAC_TRY(state_prep(c, actor, states_u, n_states, c->se_u, c->sx_u));
    c->gru_t = pack_gru(c, 0, actor, rows, /*need_v1=*/true);   // callers may pass operands this library cannot bound: round 1's kernel too
    if (!c->gru_t) {
      c->pk[PK_T_KXA] = pack_w(c, PK_T_KXA, kxa, GAC_G, GAC_E, GAC_G, 0, rows);
      c->pk[PK_T_U] = pack_w(c, PK_T_U, var(c, actor, 0, GAC_A_U), GAC_G, GAC_E, GAC_G, 0, rows);
    }
  }
  if (c->gru_t) {
    const bool bounded = (flags & GAC_GRU_ACTOR_INPUTS) != 0;
    const bool v2 = bounded && c->g2_on[0];
    if (v2) {
      // external fp32 operands: build the a16 images the v2 kernel consumes (inside the sampler the producing kernels write them)
      if (!(flags & GAC_GRU_SAME_INPUTS)) {
        GAC_CUDA(a16_pack(action_emb, GAC_E, rows, nullptr, 0, &c->g2_sc[0]->s_ae, c->ae_pk, st));
        if (h_prev) GAC_CUDA(a16_pack(h_prev, GAC_E, rows, nullptr, 0, &c->g2_sc[0]->s_h, c->h_pk[0], st));
        c->launches += h_prev ? 2 : 1;
      }
      return run_gru_step_v2(c, 0, actor, c->sx_u, state_idx, c->ae_pk, action_emb, c->h_pk[0], h_prev, h_out, c->h_pk[1], rows, nullptr);
    }
    return run_gru_step(c, 0, actor, c->sx_u, state_idx, action_emb, h_prev, h_out, rows, nullptr, nullptr, nullptr, nullptr, nullptr, bounded, nullptr);
  }
  GemmArgs g = gemm_args(rows, GAC_G, GAC_E, action_emb, GAC_E, kxa, GAC_G, c->mxa, GAC_G);
  g.Bpacked = c->pk[PK_T_KXA]; g.Bpacked_bn = c->pk_bn[PK_T_KXA];
  GAC_TRY(run_gemm(c, g));
  if (h_prev) {
    g = gemm_args(rows, GAC_G, GAC_E, h_prev, GAC_E, var(c, actor, 0, GAC_A_U), GAC_G, c->mh, GAC_G);
    g.Bpacked = c->pk[PK_T_U]; g.Bpacked_bn = c->pk_bn[PK_T_U];
    GAC_TRY(run_gemm(c, g));
  }
  gru_gates_fwd_kernel<<<(unsigned)cdiv64((long long)rows * GAC_E, 256), 256, 0, st>>>(
      c->sx_u, state_idx, c->mxa, h_prev ? c->mh : nullptr, var(c, actor, 0, GAC_A_GB) + GAC_G, h_prev, h_out, nullptr, nullptr, nullptr,
      nullptr, rows, RowValid{nullptr, 1});
  LAUNCHED(c, "gru_gates_fwd");
  return GAC_OK;
}

int gac_iqn_forward(gac_ctx_t* c, const float* actor, const float* states_u, int32_t n_states, const int32_t* state_idx,
                    const float* taus, int32_t rows, float* actions) {
  if (!c || !actor || !states_u || !state_idx || !taus || !actions || rows < 0) return fail(GAC_ERR_INVALID, "gac_iqn_forward: bad arguments");
  return iqn_forward(c, actor, states_u, n_states, state_idx, taus, rows, actions);
}

int gac_iqn_train_fwd(gac_ctx_t* c, const float* actor, const float* states_u, int32_t n_states, const int32_t* state_idx,
                      const float* taus, int32_t max_rows, const int32_t* d_rows, float* pred) {
  if (!c || !actor || !states_u || !state_idx || !taus) return fail(GAC_ERR_INVALID, "gac_iqn_train_fwd: bad arguments");
  return iqn_train_fwd(c, actor, states_u, n_states, state_idx, taus, max_rows, d_rows, pred);
}

int gac_iqn_bwd(gac_ctx_t* c, const float* actor, const float* dpred, float* grad_actor, int32_t accumulate) {
  if (!c || !actor || !dpred || !grad_actor) return fail(GAC_ERR_INVALID, "gac_iqn_bwd: bad arguments");
  return iqn_bwd(c, actor, dpred, grad_actor, accumulate);
}

int gac_critic_eval(gac_ctx_t* c, const float* fnn1, const float* fnn2, const float* states_u, const int32_t* state_idx,
                    const float* actions, int32_t rows, float* q) {
  if (!c || !fnn1 || !fnn2 || !states_u || !actions || !q) return fail(GAC_ERR_INVALID, "gac_critic_eval: bad arguments");
  return critic_eval(c, fnn1, fnn2, states_u, state_idx, actions, rows, q);
}

int gac_value_eval(gac_ctx_t* c, const float* value, const float* states, int32_t rows, float* v) {
  if (!c || !value || !states || !v) return fail(GAC_ERR_INVALID, "gac_value_eval: bad arguments");
  return value_eval(c, value, states, rows, v);
}

int gac_advantage_compact(gac_ctx_t* c, const float* q, const float* v, int32_t n, int64_t* idx_out, int32_t* pos_out, int32_t* m_out,
                          float* sum_adv_out, float* adv_out, const float* actions, int32_t A, float* actions_out,
                          const int32_t* state_idx, int32_t* sidx_out) {
  if (!c || !q || !v || !idx_out || !m_out || !sum_adv_out || n <= 0) return fail(GAC_ERR_INVALID, "gac_advantage_compact: bad arguments");
  return compact(c, q, v, n, idx_out, pos_out, m_out, sum_adv_out, adv_out, actions, A, actions_out, state_idx, sidx_out);
}

int gac_aiqn_supervised_fwd(gac_ctx_t* c, const float* actor, const float* states_u, int32_t n_states, const int32_t* state_idx,
                            const float* taus, const float* teacher, int32_t max_rows, const int32_t* d_rows, float* pred) {
  if (!c || !actor || !states_u || !state_idx || !taus || !teacher) return fail(GAC_ERR_INVALID, "gac_aiqn_supervised_fwd: bad arguments");
  return supervised_fwd(c, actor, states_u, n_states, state_idx, taus, teacher, max_rows, d_rows, pred);
}

int gac_qhuber_loss_bwd(gac_ctx_t* c, const float* pred, const float* target, const float* taus, const float* weights,
                        int32_t max_rows, const int32_t* d_rows, int32_t A, float inv_norm, float* dpred, float* loss_out) {
  if (!c || !pred || !target || !taus || !dpred || !loss_out) return fail(GAC_ERR_INVALID, "gac_qhuber_loss_bwd: bad arguments");
  return qhuber(c, pred, target, taus, weights, max_rows, d_rows, A, inv_norm, dpred, loss_out, 1);
}

int gac_aiqn_bwd(gac_ctx_t* c, const float* actor, const float* dpred, float* grad_actor, int32_t accumulate) {
  if (!c || !actor || !dpred || !grad_actor) return fail(GAC_ERR_INVALID, "gac_aiqn_bwd: bad arguments");
  return aiqn_bwd(c, actor, dpred, grad_actor, accumulate);
}

int gac_td_fwd_bwd(gac_ctx_t* c, const float* fnn, int32_t n_in, const float* x, const float* y, int32_t rows, float* grad_fnn,
                   float* loss_out) {
  if (!c || !fnn || !x || !y || !grad_fnn || !loss_out) return fail(GAC_ERR_INVALID, "gac_td_fwd_bwd: bad arguments");
  return td_fwd_bwd(c, fnn, n_in, x, y, rows, grad_fnn, loss_out);
}

int gac_replay_gather(const float* obs1, const float* obs2, const float* acts, const float* rews, const float* done, const int32_t* idx,
                      int32_t n, int32_t S, int32_t A, float* s, float* sp, float* a, float* r, float* it, void* stream) {
  if (!obs1 || !obs2 || !acts || !rews || !done || !idx || !s || !sp || !a || !r || !it || n <= 0 || S <= 0 || A <= 0)
    return fail(GAC_ERR_INVALID, "gac_replay_gather: bad arguments");
  const long long total = (long long)n * (2 * S + A + 2);
  replay_gather_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(obs1, obs2, acts, rews, done, idx, n,
                                                                                                        S, A, s, sp, a, r, it);
  GAC_LAUNCH_CHECK("replay_gather");
  return GAC_OK;
}

int gac_gemm(gac_ctx_t* c, int32_t engine, const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, int32_t transA,
             int32_t transB) {
  if (!c || !A || !B || !C) return fail(GAC_ERR_INVALID, "gac_gemm: bad arguments");
  GemmArgs g = gemm_args(M, N, K, A, transA ? M : K, B, transB ? K : N, C, N);
  g.transA = transA; g.transB = transB;
  if (engine == 0 || engine == 2) {
    if (!tc_gemm_supported(g)) return fail(GAC_ERR_UNSUPPORTED, "gac_gemm: shape not supported by the tcgen05 engine");
    if (engine == 2) {   // B treated as a weight matrix: pre-split / pre-swizzled once, fetched by bulk copies
      if (transA || tc_packed_bytes(K, N) > c->pack_bytes) return fail(GAC_ERR_UNSUPPORTED, "gac_gemm: packed-B needs !transA and a weight-sized B");
      const int pbn = (tc_pair_enabled() && M >= 1024) ? tc::pair_bn(N) : 0;
      GAC_CUDA(tc_pack_b(B, g.ldb, transB, K, N, c->pack[15], c->stream, pbn));
      c->launches++;
      g.Bpacked = c->pack[15]; g.Bpacked_bn = pbn;
      if (getenv("GAC_TC_TIMELINE")) g.dbg = reinterpret_cast<unsigned long long*>(c->spart);   // scratch: 16 x 8 stamps
    }
    GAC_CUDA(launch_gemm_tc(g, c->stream));
  } else {
    GAC_CUDA(launch_gemm_simt(g, c->stream));
  }
  c->launches++;
  return GAC_OK;
}

// ---- the train step -------------------------------------------------------------------------
int gac_step_grads(gac_ctx_t* c, const gac_step_io_t* io, const gac_hparams_t* hp) {
  if (!c || !io || !hp) return fail(GAC_ERR_INVALID, "gac_step_grads: null argument");
  if (hp->mode != 0 && hp->mode != 1) return fail(GAC_ERR_UNSUPPORTED, "target density mode must be linear or boltzmann (networks.py:94-95)");
  const int S = c->cfg.S, A = c->cfg.A, B = c->cfg.B, K = c->cfg.K, P = c->P, Mc = c->Mc;
  cudaStream_t st = c->stream;
  const gac_layout_t& L = c->lay;
  const float* t_actor = io->target + L.net_begin[GAC_NET_ACTOR];
  const float* t_f1 = io->target + L.net_begin[GAC_NET_FNN1];
  const float* t_f2 = io->target + L.net_begin[GAC_NET_FNN2];
  const float* t_v = io->target + L.net_begin[GAC_NET_VALUE];
  float* stats = io->grad + L.net_begin[GAC_NNETS];

  // row -> state maps: [0,P) state-major (networks.py:464-467), [P,3P) sample-major (agent.py:186)
  // (static per (B,K); rebuilt each step: 3P ints)
  sidx3_kernel<<<cdiv(3 * P, 256), 256, 0, st>>>(c->sidx3, P, B, K);
  LAUNCHED(c, "sidx3");

  // Two lanes.  The three TD updates are B-row problems (~65 launches of a few microseconds each) whose results nobody reads
  // before gac_step_apply: they run on the side stream, with their own split-K / column-sum scratch, in the gaps the big
  // kernels of the main lane leave (launch tails, partial waves).  Fork / join by events, so a CUDA-graph capture of the
  // caller's stream takes the side lane along.  GAC_NO_SIDE_STREAM=1: everything on the caller's stream, in program order.
  static const bool no_side = getenv("GAC_NO_SIDE_STREAM") != nullptr;
  const bool two_lanes = !no_side && c->side != nullptr;
  float* const spart_main = c->spart; float* const cpart_main = c->cpart;
  struct LaneGuard {                                              // whatever path leaves this function, the context is back on the main lane
    gac_ctx* c; cudaStream_t st; float *sp, *cp;
    ~LaneGuard() { c->stream = st; c->spart = sp; c->cpart = cp; c->lanes_on = false; c->critic_prepped = false; c->actor_prep_token = nullptr; }
  } lane_guard{c, st, spart_main, cpart_main};
  c->lanes_on = two_lanes; c->main_stream = st; c->spart_main = spart_main; c->cpart_main = cpart_main;
  auto side_lane = [&]() { lane_side(c); };
  auto main_lane = [&]() { lane_main(c); };
  if (two_lanes) {
    GAC_CUDA(cudaEventRecord(c->ev_fork, st));                    // the batch and the parameters are ready at this point of the caller's stream
    GAC_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
  }

  // weight-only preparation of the critic evaluation and of the actor's training pass: first on the side lane, so that it is
  // long done when the main lane gets there (and done ONCE for all row chunks of the actor update)
  if (two_lanes) {
    side_lane();
    if (critic_fused(c, 3 * P, B, c->sidx3)) {
      GAC_TRY(critic_prep(c, t_f1, t_f2, io->s, B));
      c->critic_prepped = true;
    }
    GAC_CUDA(cudaEventRecord(c->ev_cprep, c->side));
    if (c->cfg.actor_kind == GAC_ACTOR_AIQN) {
      GAC_TRY(actor_train_prep(c, io->online + L.net_begin[GAC_NET_ACTOR]));
      c->actor_prep_token = io->online + L.net_begin[GAC_NET_ACTOR];
    }
    GAC_CUDA(cudaEventRecord(c->ev_aprep, c->side));
    main_lane();
  }

  // ---- Critic.train (networks.py:410-426) ----
  side_lane();
  critic_train_x_kernel<<<cdiv(B * (S + A), 256), 256, 0, c->stream>>>(io->s, io->a, io->critic_u, hp->q_normalization, S, A, B, c->td_x);
  LAUNCHED(c, "critic_train_x");
  GAC_TRY(value_eval(c, t_v, io->sp, B, c->td_vnext, c->td_h0, c->td_h1));
  td_target_kernel<<<cdiv(B, 256), 256, 0, c->stream>>>(io->r, c->td_vnext, io->done, hp->gamma, B, c->td_y);
  LAUNCHED(c, "td_target");
  GAC_TRY(td_fwd_bwd(c, io->online + L.net_begin[GAC_NET_FNN1], S + A, c->td_x, c->td_y, B, io->grad + L.net_begin[GAC_NET_FNN1], stats + 0));
  GAC_TRY(td_fwd_bwd(c, io->online + L.net_begin[GAC_NET_FNN2], S + A, c->td_x, c->td_y, B, io->grad + L.net_begin[GAC_NET_FNN2], stats + 1));
  main_lane();

  // ---- both target-actor sampling passes in one 2P-row batch (F9) ----
  GAC_CUDA(cudaMemcpyAsync(c->taus2, io->value_tau, (size_t)P * A * sizeof(float), cudaMemcpyDeviceToDevice, st));
  GAC_CUDA(cudaMemcpyAsync(c->taus2 + (size_t)P * A, io->filter_tau, (size_t)P * A * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (c->cfg.actor_kind == GAC_ACTOR_IQN) GAC_TRY(iqn_forward(c, t_actor, io->s, B, c->sidx3, c->taus2, 2 * P, c->acts2));
  else GAC_TRY(aiqn_sample(c, t_actor, io->s, B, c->sidx3, c->taus2, 2 * P, c->acts2));
  // candidates: [value-pass actions ; clip(target actions + 0.01 N) ; uniform random actions]
  GAC_CUDA(cudaMemcpyAsync(c->cand, c->acts2, (size_t)P * A * sizeof(float), cudaMemcpyDeviceToDevice, st));
  noise_clip_kernel<<<(unsigned)cdiv64((long long)P * A, 256), 256, 0, st>>>(c->acts2 + (size_t)P * A, io->filter_normal, 0.01f,
                                                                           (long long)P * A, c->cand + (size_t)P * A);
  LAUNCHED(c, "noise_clip");
  GAC_CUDA(cudaMemcpyAsync(c->cand + (size_t)2 * P * A, io->filter_rand, (size_t)P * A * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (two_lanes) GAC_CUDA(cudaStreamWaitEvent(st, c->ev_cprep, 0));
  GAC_TRY(critic_eval(c, t_f1, t_f2, io->s, c->sidx3, c->cand, 3 * P, c->q3, B));

  // ---- Value.train (networks.py:483-496): needs the critics' values of the first P candidates ----
  if (two_lanes) {
    GAC_CUDA(cudaEventRecord(c->ev_q, st));
    GAC_CUDA(cudaStreamWaitEvent(c->side, c->ev_q, 0));
  }
  side_lane();
  mean_k_kernel<<<cdiv(B * 32, 256), 256, 0, c->stream>>>(c->q3, K, B, c->td_vcrit);
  LAUNCHED(c, "mean_k");
  GAC_TRY(td_fwd_bwd(c, io->online + L.net_begin[GAC_NET_VALUE], S, io->s, c->td_vcrit, B, io->grad + L.net_begin[GAC_NET_VALUE], stats + 2));
  if (two_lanes) GAC_CUDA(cudaEventRecord(c->ev_join, c->side));
  main_lane();

  // ---- advantage filter (agent.py:198-215) ----
  GAC_TRY(value_eval(c, t_v, io->s, B, c->vt));
  gather_kernel<<<cdiv(2 * P, 256), 256, 0, st>>>(c->vt, c->sidx3 + P, 2 * P, c->v2);
  LAUNCHED(c, "gather_v");
  GAC_TRY(compact(c, c->q3 + P, c->v2, 2 * P, c->idx_sel, c->pos, c->m_dev, c->sum_adv, c->adv_sel, c->cand + (size_t)P * A, A,
                  c->act_sel, c->sidx3 + P, c->sidx_sel));
  chunk_rows_kernel<<<cdiv(c->nchunks, 1024), 1024, 0, st>>>(c->m_dev, Mc, c->nchunks, c->rows_c);
  LAUNCHED(c, "chunk_rows");
  stats_m_kernel<<<1, 32, 0, st>>>(c->m_dev, stats);
  LAUNCHED(c, "stats_m");
  GAC_CUDA(cudaMemcpyAsync(stats + 4, c->sum_adv, sizeof(float), cudaMemcpyDeviceToDevice, st));
  GAC_CUDA(cudaMemsetAsync(stats + 3, 0, sizeof(float), st));

  // ---- IQNActor.train (networks.py:122-141) in row chunks; un-normalised weights ----
  float* g_actor = io->grad + L.net_begin[GAC_NET_ACTOR];
  const float* o_actor = io->online + L.net_begin[GAC_NET_ACTOR];
  if (two_lanes) GAC_CUDA(cudaStreamWaitEvent(st, c->ev_aprep, 0));
  for (int ch = 0; ch < c->nchunks; ++ch) {
    const size_t r0 = (size_t)ch * Mc;
    const int maxr = (int)((size_t)2 * P - r0 < (size_t)Mc ? (size_t)2 * P - r0 : (size_t)Mc);
    if (c->cfg.actor_kind == GAC_ACTOR_IQN)
      GAC_TRY(iqn_train_fwd(c, o_actor, io->s, B, c->sidx_sel + r0, io->actor_tau + r0 * A, maxr, c->rows_c + ch, nullptr));
    else
      GAC_TRY(supervised_fwd(c, o_actor, io->s, B, c->sidx_sel + r0, io->actor_tau + r0 * A, c->act_sel + r0 * A, maxr, c->rows_c + ch, nullptr));
    GAC_TRY(qhuber(c, c->pred_c, c->act_sel + r0 * A, io->actor_tau + r0 * A, hp->mode == 0 ? c->adv_sel + r0 : nullptr, maxr,
                   c->rows_c + ch, A, 1.f, c->dpred_c, stats + 3, 1));
    if (c->cfg.actor_kind == GAC_ACTOR_IQN) GAC_TRY(iqn_bwd(c, o_actor, c->dpred_c, g_actor, ch > 0));
    else GAC_TRY(aiqn_bwd(c, o_actor, c->dpred_c, g_actor, ch > 0));
  }

  if (two_lanes) GAC_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0));   // join: gac_step_apply (and the caller's all-reduce) read the TD gradients
  // optional taps for the parity tests
  if (io->sel_idx) GAC_CUDA(cudaMemcpyAsync(io->sel_idx, c->idx_sel, (size_t)2 * P * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  if (io->q_out) GAC_CUDA(cudaMemcpyAsync(io->q_out, c->q3 + P, (size_t)2 * P * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (io->v_out) GAC_CUDA(cudaMemcpyAsync(io->v_out, c->v2, (size_t)2 * P * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (io->cand_actions) GAC_CUDA(cudaMemcpyAsync(io->cand_actions, c->cand + (size_t)P * A, (size_t)2 * P * A * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (io->value_actions) GAC_CUDA(cudaMemcpyAsync(io->value_actions, c->cand, (size_t)P * A * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return GAC_OK;
}

int gac_step_apply(gac_ctx_t* c, const gac_step_io_t* io, const gac_hparams_t* hp) {
  if (!c || !io || !hp) return fail(GAC_ERR_INVALID, "gac_step_apply: null argument");
  cudaStream_t st = c->stream;
  const gac_layout_t& L = c->lay;
  const float* stats = io->grad + L.net_begin[GAC_NNETS];
  apply_prep_kernel<<<1, 32, 0, st>>>(stats, c->cfg.A, hp->mode, c->gscale, c->skip);
  LAUNCHED(c, "apply_prep");
  return gac_adam_polyak(c, io->online, io->target, io->adam_m, io->adam_v, io->grad, hp, io->adam_t, c->gscale, c->skip);
}

}  // extern "C"
