/* TEST INFRASTRUCTURE / BASELINE ONLY — whole-token CPU decode of a Llama-family GGUF model with the
 * reference's CPU arithmetic (the candle / llama.cpp QMatMul path restated in mrs_oracle.c: Q8_K /
 * Q8_0 activation quantisation + integer block dots, REF mistralrs-quant/src/gguf/mod.rs:472
 * `QMatMul::forward`, candle-core quantized::k_quants vec_dot), timed on the box's host cores:
 *
 *   - ALL layers of the model (every weight byte of a token is streamed from DRAM: the weights, e.g.
 *     4.6 GB for Llama-3-8B Q4_K_M, are far larger than any cache, no layer is sampled or repeated);
 *   - RMSNorm, RoPE-free attention over a `ctx`-token f32 KV history (per-head softmax), SiLU*mul and
 *     the residual adds are included, parallel over heads / rows;
 *   - one persistent thread per host core, pinned, spin barriers between the 7 GEMVs of a layer (the
 *     reference's rayon pool plays this role), row-range work split;
 *   - compiled on the box at run time with -O3 -march=native (bench.py) — never shipped prebuilt.
 *
 * Weights are synthetic bytes (the dot kernels' cost does not depend on the values); block scales are
 * patched to small finite f16 numbers so no NaN/Inf slow paths are hit.  Parity of the arithmetic is
 * the oracle's business (tests/), this file only times it. */
#define _GNU_SOURCE
#include "mrs_oracle.c"

#include <sched.h>
#include <stdatomic.h>
#include <time.h>

typedef struct { int type, rows, cols; uint8_t *w; } cpu_mat_t;
typedef struct { cpu_mat_t q, k, v, o, gate, up, down; float *attn_norm, *ffn_norm; float *kcache, *vcache; } cpu_layer_t;

static struct {
  int nthreads, hidden, inter, nq, nkv, heads, kv_heads, head_dim, vocab, n_layers, ctx;
  cpu_layer_t *layers; cpu_mat_t lm_head; float *final_norm;
  /* activations */
  float *x, *h, *q, *k, *v, *attn, *o, *gate, *up, *act, *logits;
  void *yq; float *yd;
  /* job */
  atomic_int phase; atomic_int arrived; atomic_int stop;
  int job_kind; const cpu_mat_t *job_mat; const float *job_in; float *job_out; int job_layer;
} G;

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint64_t xorshift(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

static void fill_mat(cpu_mat_t *m, int type, int rows, int cols) {
  const int be = mrs_block_elems(type), bb = mrs_block_bytes(type);
  const size_t nblk = (size_t)rows * cols / be, bytes = nblk * bb;
  m->type = type; m->rows = rows; m->cols = cols;
  m->w = (uint8_t *)malloc(bytes + 64);
  uint64_t *p = (uint64_t *)m->w;
  for (size_t i = 0; i < (bytes + 7) / 8; i++) p[i] = xorshift();
  /* finite small f16 scales: 0x1c00 = 2^-8 */
  for (size_t b = 0; b < nblk; b++) {
    uint8_t *blk = m->w + b * bb;
    int off1 = 0, off2 = -1;
    switch (type) {
    case MRS_Q4_K: case MRS_Q5_K: off1 = 0; off2 = 2; break;
    case MRS_Q6_K: off1 = 208; break;
    case MRS_Q2_K: off1 = 80; off2 = 82; break;
    case MRS_Q3_K: off1 = 108; break;
    case MRS_Q4_1: case MRS_Q5_1: off1 = 0; off2 = 2; break;
    default: off1 = 0; break;
    }
    blk[off1] = 0x00; blk[off1 + 1] = 0x1c;
    if (off2 >= 0) { blk[off2] = 0x00; blk[off2 + 1] = 0x18; }
  }
}

static void gemv_rows(const cpu_mat_t *m, const float *unused, float *out, int r0, int r1) {
  (void)unused;
  cpu_job_t job = {m->type, m->cols, r0, r1, m->rows, 1, m->w, G.yq, G.yd, out};
  cpu_job_run(&job);
}

/* attention of heads [h0, h1) over the ctx-token history + the new token (f32, per-head softmax) */
static void attn_heads(int layer, int h0, int h1) {
  const cpu_layer_t *L = &G.layers[layer];
  const int D = G.head_dim, g = G.heads / G.kv_heads, T = G.ctx;
  const float scale = 1.0f / sqrtf((float)D);
  float *s = (float *)alloca(sizeof(float) * (T + 1));
  for (int h = h0; h < h1; h++) {
    const int kvh = h / g;
    const float *qh = G.q + (size_t)h * D;
    float mx = -INFINITY;
    for (int t = 0; t <= T; t++) {
      const float *kt = (t < T) ? L->kcache + ((size_t)t * G.kv_heads + kvh) * D : G.k + (size_t)kvh * D;
      float a = 0.f;
      for (int d = 0; d < D; d++) a += qh[d] * kt[d];
      s[t] = a * scale; if (s[t] > mx) mx = s[t];
    }
    float sum = 0.f;
    for (int t = 0; t <= T; t++) { s[t] = expf(s[t] - mx); sum += s[t]; }
    float *oh = G.attn + (size_t)h * D;
    for (int d = 0; d < D; d++) oh[d] = 0.f;
    for (int t = 0; t <= T; t++) {
      const float *vt = (t < T) ? L->vcache + ((size_t)t * G.kv_heads + kvh) * D : G.v + (size_t)kvh * D;
      const float p = s[t] / sum;
      for (int d = 0; d < D; d++) oh[d] += p * vt[d];
    }
  }
}

enum { JOB_GEMV = 1, JOB_ATTN = 2, JOB_GLU = 3 };
#define MRS_SPIN_LIMIT 20000

static void run_share(int tid) {
  const int nt = G.nthreads;
  if (G.job_kind == JOB_GEMV) {
    const int rows = G.job_mat->rows;
    gemv_rows(G.job_mat, G.job_in, G.job_out, (int)((int64_t)rows * tid / nt), (int)((int64_t)rows * (tid + 1) / nt));
  } else if (G.job_kind == JOB_ATTN) {
    attn_heads(G.job_layer, (int)((int64_t)G.heads * tid / nt), (int)((int64_t)G.heads * (tid + 1) / nt));
  } else if (G.job_kind == JOB_GLU) {
    const int n = G.inter, i0 = (int)((int64_t)n * tid / nt), i1 = (int)((int64_t)n * (tid + 1) / nt);
    for (int i = i0; i < i1; i++) G.act[i] = G.gate[i] / (1.0f + expf(-G.gate[i])) * G.up[i];
  }
}

static void pin_to(int idx) {
  cpu_set_t all, got;
  CPU_ZERO(&all);
  for (int c = 0; c < CPU_SETSIZE; c++) CPU_SET(c, &all);
  pthread_setaffinity_np(pthread_self(), sizeof all, &all);
  if (pthread_getaffinity_np(pthread_self(), sizeof got, &got) != 0) return;
  const int n = CPU_COUNT(&got);
  if (n <= 1) return;
  int want = idx % n, k = 0;
  for (int c = 0; c < CPU_SETSIZE; c++) {
    if (!CPU_ISSET(c, &got)) continue;
    if (k++ == want) { cpu_set_t one; CPU_ZERO(&one); CPU_SET(c, &one); pthread_setaffinity_np(pthread_self(), sizeof one, &one); break; }
  }
}

static void *worker(void *arg) {
  const int tid = (int)(intptr_t)arg;
  pin_to(tid);
  int seen = 0;
  for (;;) {
    for (int spins = 0; atomic_load_explicit(&G.phase, memory_order_acquire) == seen; spins++) {
      if (atomic_load_explicit(&G.stop, memory_order_relaxed)) return NULL;
      if (spins < MRS_SPIN_LIMIT) MRS_CPU_RELAX(); else sched_yield();   /* a CPU quota below the thread count must not starve the pool */
    }
    seen = atomic_load_explicit(&G.phase, memory_order_acquire);
    run_share(tid);
    atomic_fetch_add_explicit(&G.arrived, 1, memory_order_release);
  }
}

static void dispatch(int kind, const cpu_mat_t *m, const float *in, float *out, int layer) {
  G.job_kind = kind; G.job_mat = m; G.job_in = in; G.job_out = out; G.job_layer = layer;
  atomic_store_explicit(&G.arrived, 0, memory_order_relaxed);
  atomic_fetch_add_explicit(&G.phase, 1, memory_order_release);
  run_share(0);
  for (int spins = 0; atomic_load_explicit(&G.arrived, memory_order_acquire) < G.nthreads - 1; spins++) {
    if (spins < MRS_SPIN_LIMIT) MRS_CPU_RELAX(); else sched_yield();
  }
}

static void quantize_act(const float *x, int cols, int be) {
  if (be == 256) quantize_row_q8_k(x, (blk_q8_k *)G.yq, cols);
  else quantize_row_q8_0(x, (blk_q8_0 *)G.yq, G.yd, cols);
}
static void gemv(const cpu_mat_t *m, const float *in, float *out) { dispatch(JOB_GEMV, m, in, out, 0); }

static void rmsnorm(const float *x, const float *w, float *out, int n) {
  double ss = 0;
  for (int i = 0; i < n; i++) ss += (double)x[i] * x[i];
  const float inv = 1.0f / sqrtf((float)(ss / n) + 1e-5f);
  for (int i = 0; i < n; i++) out[i] = x[i] * inv * w[i];
}

static void one_token(void) {
  const int H = G.hidden;
  for (int l = 0; l < G.n_layers; l++) {
    cpu_layer_t *L = &G.layers[l];
    rmsnorm(G.x, L->attn_norm, G.h, H);
    quantize_act(G.h, H, mrs_block_elems(L->q.type));
    gemv(&L->q, G.h, G.q); gemv(&L->k, G.h, G.k);
    if (mrs_block_elems(L->v.type) != mrs_block_elems(L->q.type)) quantize_act(G.h, H, mrs_block_elems(L->v.type));
    gemv(&L->v, G.h, G.v);
    dispatch(JOB_ATTN, NULL, NULL, NULL, l);
    quantize_act(G.attn, G.nq, mrs_block_elems(L->o.type));
    gemv(&L->o, G.attn, G.o);
    for (int i = 0; i < H; i++) G.x[i] += G.o[i];
    rmsnorm(G.x, L->ffn_norm, G.h, H);
    quantize_act(G.h, H, mrs_block_elems(L->gate.type));
    gemv(&L->gate, G.h, G.gate); gemv(&L->up, G.h, G.up);
    dispatch(JOB_GLU, NULL, NULL, NULL, l);
    quantize_act(G.act, G.inter, mrs_block_elems(L->down.type));
    gemv(&L->down, G.act, G.o);
    for (int i = 0; i < H; i++) G.x[i] = 0.5f * (G.x[i] + G.o[i]);   /* keep the synthetic stream bounded */
  }
  rmsnorm(G.x, G.final_norm, G.h, H);
  quantize_act(G.h, H, mrs_block_elems(G.lm_head.type));
  gemv(&G.lm_head, G.h, G.logits);
}

static float *randf(size_t n, float scale) {
  float *p = (float *)malloc(n * sizeof(float));
  for (size_t i = 0; i < n; i++) p[i] = scale * ((float)(xorshift() >> 40) / 8388608.0f - 1.0f);
  return p;
}

/* types: per layer [q, k, v, o, gate, up, down] ggml codes (7 * n_layers ints) + lm_head type.
 * Runs whole tokens for about `seconds` (at least `min_tokens`); returns tokens/s, fills *tokens_run,
 * *weight_bytes.  threads <= 0: one per allowed CPU. */
/* nsamples timed samples of about `seconds` each on ONE model allocation (the weights are generated and paged in once):
 * tok_s[i] and tokens[i] for sample i; returns the mean tokens/s. */
double mrs_cpu_decode_bench_samples(int n_layers, int hidden, int inter, int heads, int kv_heads, int head_dim, int vocab,
                                    const int *types, int lm_head_type, int ctx, int threads, double seconds, int min_tokens,
                                    int nsamples, double *tok_s, int *tokens, double *weight_bytes, int *threads_used) {
  memset(&G, 0, sizeof G);
  if (threads <= 0) { cpu_set_t got; threads = (sched_getaffinity(0, sizeof got, &got) == 0) ? CPU_COUNT(&got) : 1; }
  if (threads > MRS_MAX_THREADS) threads = MRS_MAX_THREADS;
  G.nthreads = threads; G.hidden = hidden; G.inter = inter; G.heads = heads; G.kv_heads = kv_heads; G.head_dim = head_dim;
  G.nq = heads * head_dim; G.nkv = kv_heads * head_dim; G.vocab = vocab; G.n_layers = n_layers; G.ctx = ctx;
  G.layers = (cpu_layer_t *)calloc(n_layers, sizeof(cpu_layer_t));
  double wb = 0;
  for (int l = 0; l < n_layers; l++) {
    cpu_layer_t *L = &G.layers[l];
    const int *t = types + 7 * l;
    fill_mat(&L->q, t[0], G.nq, hidden); fill_mat(&L->k, t[1], G.nkv, hidden); fill_mat(&L->v, t[2], G.nkv, hidden);
    fill_mat(&L->o, t[3], hidden, G.nq); fill_mat(&L->gate, t[4], inter, hidden); fill_mat(&L->up, t[5], inter, hidden);
    fill_mat(&L->down, t[6], hidden, inter);
    const cpu_mat_t *ms[7] = {&L->q, &L->k, &L->v, &L->o, &L->gate, &L->up, &L->down};
    for (int i = 0; i < 7; i++) wb += (double)ms[i]->rows * ms[i]->cols / mrs_block_elems(ms[i]->type) * mrs_block_bytes(ms[i]->type);
    L->attn_norm = randf(hidden, 1.0f); L->ffn_norm = randf(hidden, 1.0f);
    L->kcache = randf((size_t)ctx * G.nkv, 1.0f); L->vcache = randf((size_t)ctx * G.nkv, 1.0f);
  }
  fill_mat(&G.lm_head, lm_head_type, vocab, hidden);
  wb += (double)vocab * hidden / mrs_block_elems(lm_head_type) * mrs_block_bytes(lm_head_type);
  G.final_norm = randf(hidden, 1.0f);
  G.x = randf(hidden, 1.0f); G.h = randf(hidden, 1.0f); G.q = randf(G.nq, 1.0f); G.k = randf(G.nkv, 1.0f); G.v = randf(G.nkv, 1.0f);
  G.attn = randf(G.nq, 1.0f); G.o = randf(hidden, 1.0f); G.gate = randf(inter, 1.0f); G.up = randf(inter, 1.0f);
  G.act = randf(inter, 1.0f); G.logits = randf(vocab, 1.0f);
  const int maxk = inter > G.nq ? inter : G.nq;
  G.yq = malloc(sizeof(blk_q8_k) * (size_t)(maxk / 32 + 8));
  G.yd = (float *)malloc(sizeof(float) * (size_t)(maxk / 32 + 8));
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
  pin_to(0);
  for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, worker, (void *)(intptr_t)t);
  one_token();                              /* warm-up: page the weights in, wake the pool */
  double mean = 0;
  for (int s = 0; s < nsamples; s++) {
    int n = 0;
    const double t0 = now_s();
    while (n < min_tokens || now_s() - t0 < seconds) { one_token(); n++; }
    const double dt = now_s() - t0;
    if (tok_s) tok_s[s] = n / dt;
    if (tokens) tokens[s] = n;
    mean += n / dt / nsamples;
  }
  atomic_store(&G.stop, 1);
  for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
  if (weight_bytes) *weight_bytes = wb;
  if (threads_used) *threads_used = threads;
  /* (memory is released at process exit: the bench runs this once in a short-lived child) */
  return mean;
}

double mrs_cpu_decode_bench(int n_layers, int hidden, int inter, int heads, int kv_heads, int head_dim, int vocab,
                            const int *types, int lm_head_type, int ctx, int threads, double seconds, int min_tokens,
                            int *tokens_run, double *weight_bytes, int *threads_used) {
  double ts = 0;
  int n = 0;
  mrs_cpu_decode_bench_samples(n_layers, hidden, inter, heads, kv_heads, head_dim, vocab, types, lm_head_type, ctx, threads, seconds,
                               min_tokens, 1, &ts, &n, weight_bytes, threads_used);
  if (tokens_run) *tokens_run = n;
  return ts;
}
