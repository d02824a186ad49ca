/*
 * oracle_ref.c -- plain-C CPU restatement of the ark-crypto-primitives hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under crypto_primitives_b200/ or include/ may
 * link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do, as the checker and the timed CPU baseline.
 *
 * It is a RESTATEMENT ("port"), not the reference: the reference is Rust and its
 * arithmetic lives in ark-ff / ark-ec ^0.4, which are not under /root/reference and
 * cannot be built here (no rustc/cargo in this image).  Each function cites the
 * reference lines it follows (R = /root/reference/crypto-primitives/src).  It keeps
 * the reference's algorithmic shape on purpose -- dense t x t MDS every round,
 * generic square-and-multiply x^alpha, one hash at a time, per-level barriers --
 * because it doubles as the timed "C restatement of the reference CPU path".
 * Data parallelism mirrors the reference's rayon `parallel` feature
 * (R/merkle_tree/mod.rs:417,458,494) with pthreads.
 *
 * Interchange: field elements are 4 x u64 little-endian limbs in Montgomery form,
 * R = 2^256, fully reduced (what ark-ff's Fp<MontBackend<_,4>,4> holds).
 *
 * Pinning: Poseidon results are pinned by the reference KATs through the Python
 * oracle (tests/test_oracle_*.py check C == Python == KAT).  Pedersen / Merkle:
 * PARITY UNPINNED (no golden vectors exist in the reference).
 */
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

typedef struct { u64 l[4]; } fe;

typedef struct {
    u64 p[4];
    u64 ninv;      /* -p^{-1} mod 2^64 */
    fe one;        /* R mod p */
    fe r2;         /* R^2 mod p */
    int nocarry;   /* top bit of p clear: fe_mul may drop the carry word */
} field_t;

/* ---------------------------------------------------------------- field */

static inline int ge4(const u64 a[4], const u64 b[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}

static inline u64 sub4(u64 r[4], const u64 a[4], const u64 b[4]) {
    unsigned long long t;
    unsigned char bw = _subborrow_u64(0, a[0], b[0], &t); r[0] = t;
    bw = _subborrow_u64(bw, a[1], b[1], &t); r[1] = t;
    bw = _subborrow_u64(bw, a[2], b[2], &t); r[2] = t;
    bw = _subborrow_u64(bw, a[3], b[3], &t); r[3] = t;
    return bw;
}

static inline u64 add4(u64 r[4], const u64 a[4], const u64 b[4]) {
    unsigned long long t;
    unsigned char c = _addcarry_u64(0, a[0], b[0], &t); r[0] = t;
    c = _addcarry_u64(c, a[1], b[1], &t); r[1] = t;
    c = _addcarry_u64(c, a[2], b[2], &t); r[2] = t;
    c = _addcarry_u64(c, a[3], b[3], &t); r[3] = t;
    return c;
}

/* a + b mod p, a - b mod p: add/sub then one conditional correction, branch-free (what ark-ff's add_assign /
 * sub_assign + subtract_modulus do with a data-dependent branch; same values). */
static inline void fe_add(const field_t *F, fe *r, const fe *a, const fe *b) {
    u64 t[4], d[4];
    u64 c = add4(t, a->l, b->l);
    u64 bw = sub4(d, t, F->p);
    int keep = !c && bw;                      /* t < p: keep t */
    for (int i = 0; i < 4; i++) r->l[i] = keep ? t[i] : d[i];
}

static inline void fe_sub(const field_t *F, fe *r, const fe *a, const fe *b) {
    u64 t[4], d[4];
    u64 bw = sub4(t, a->l, b->l);
    add4(d, t, F->p);
    for (int i = 0; i < 4; i++) r->l[i] = bw ? d[i] : t[i];
}

/* Montgomery product a*b/R mod p, 64-bit limbs, fully unrolled.  Two forms, both standard:
 *  - fe_mul_nocarry: CIOS without the extra carry word, valid when the modulus' top bit is clear (every field of this
 *    path; the same shortcut ark-ff's MontBackend takes for such moduli -- dep, from memory): per row
 *    (A,t0) = t0 + a0*bi; m = t0*ninv; (C,_) = t0 + m*p0; (A,tj) = tj + aj*bi + A; (C,t(j-1)) = tj + m*pj + C; t3 = C + A.
 *  - fe_mul_general: textbook CIOS with the carry word, any modulus < 2^256.
 * The compiler keeps t[] in registers and emits mulx/adc code with -march=x86-64-v3. */
#define MAC(acc, x, y, carry)                                  \
    do {                                                       \
        u128 s__ = (u128)(x) * (y) + (acc) + (carry);          \
        (acc) = (u64)s__;                                      \
        (carry) = (u64)(s__ >> 64);                            \
    } while (0)
static inline void fe_mul_general(const field_t *F, fe *r, const fe *a, const fe *b) {
    const u64 p0 = F->p[0], p1 = F->p[1], p2 = F->p[2], p3 = F->p[3], ninv = F->ninv;
    const u64 a0 = a->l[0], a1 = a->l[1], a2 = a->l[2], a3 = a->l[3];
    u64 t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    for (int i = 0; i < 4; i++) {
        const u64 bi = b->l[i];
        u64 c = 0, t5;
        MAC(t0, a0, bi, c); MAC(t1, a1, bi, c); MAC(t2, a2, bi, c); MAC(t3, a3, bi, c);
        { u128 s = (u128)t4 + c; t4 = (u64)s; t5 = (u64)(s >> 64); }
        const u64 m = t0 * ninv;
        c = 0;
        { u64 lo = t0; MAC(lo, m, p0, c); }
        { u64 v = t1; MAC(v, m, p1, c); t0 = v; }
        { u64 v = t2; MAC(v, m, p2, c); t1 = v; }
        { u64 v = t3; MAC(v, m, p3, c); t2 = v; }
        { u128 s = (u128)t4 + c; t3 = (u64)s; t4 = t5 + (u64)(s >> 64); }
    }
    u64 t[4] = {t0, t1, t2, t3};
    if (t4 || ge4(t, F->p)) sub4(t, t, F->p);
    memcpy(r->l, t, 32);
}

#define NC_ROW(bi)                                                                      \
    do {                                                                                \
        u128 s__, q__;                                                                  \
        u64 A__, C__, m__;                                                              \
        s__ = (u128)a0 * (bi) + t0;                A__ = (u64)(s__ >> 64);              \
        m__ = (u64)s__ * ninv;                                                          \
        q__ = (u128)m__ * p0 + (u64)s__;           C__ = (u64)(q__ >> 64);              \
        s__ = (u128)a1 * (bi) + t1 + A__;          A__ = (u64)(s__ >> 64);              \
        q__ = (u128)m__ * p1 + (u64)s__ + C__;     t0 = (u64)q__; C__ = (u64)(q__ >> 64); \
        s__ = (u128)a2 * (bi) + t2 + A__;          A__ = (u64)(s__ >> 64);              \
        q__ = (u128)m__ * p2 + (u64)s__ + C__;     t1 = (u64)q__; C__ = (u64)(q__ >> 64); \
        s__ = (u128)a3 * (bi) + t3 + A__;          A__ = (u64)(s__ >> 64);              \
        q__ = (u128)m__ * p3 + (u64)s__ + C__;     t2 = (u64)q__; C__ = (u64)(q__ >> 64); \
        t3 = C__ + A__;                                                                 \
    } while (0)
static inline void fe_mul_nocarry(const field_t *F, fe *r, const fe *a, const fe *b) {
    const u64 p0 = F->p[0], p1 = F->p[1], p2 = F->p[2], p3 = F->p[3], ninv = F->ninv;
    const u64 a0 = a->l[0], a1 = a->l[1], a2 = a->l[2], a3 = a->l[3];
    const u64 b0 = b->l[0], b1 = b->l[1], b2 = b->l[2], b3 = b->l[3];
    u64 t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    NC_ROW(b0); NC_ROW(b1); NC_ROW(b2); NC_ROW(b3);
    /* result < 2p: one conditional subtraction, branch-free */
    u64 d0, d1, d2, d3, bw;
    { u128 d = (u128)t0 - p0;      d0 = (u64)d; bw = (u64)(d >> 64) & 1; }
    { u128 d = (u128)t1 - p1 - bw; d1 = (u64)d; bw = (u64)(d >> 64) & 1; }
    { u128 d = (u128)t2 - p2 - bw; d2 = (u64)d; bw = (u64)(d >> 64) & 1; }
    { u128 d = (u128)t3 - p3 - bw; d3 = (u64)d; bw = (u64)(d >> 64) & 1; }
    r->l[0] = bw ? t0 : d0; r->l[1] = bw ? t1 : d1; r->l[2] = bw ? t2 : d2; r->l[3] = bw ? t3 : d3;
}
static inline void fe_mul(const field_t *F, fe *r, const fe *a, const fe *b) {
    if (F->nocarry) fe_mul_nocarry(F, r, a, b);
    else fe_mul_general(F, r, a, b);
}

/* a^2/R mod p: the six cross products once, doubled, plus the four diagonal squares, then four Montgomery
 * reduction rows over the 8-limb square (ark-ff's square_in_place has the same shape -- dep, from memory). */
static inline void fe_sqr(const field_t *F, fe *r, const fe *a) {
    if (!F->nocarry) { fe_mul_general(F, r, a, a); return; }
    const u64 p0 = F->p[0], p1 = F->p[1], p2 = F->p[2], p3 = F->p[3], ninv = F->ninv;
    const u64 a0 = a->l[0], a1 = a->l[1], a2 = a->l[2], a3 = a->l[3];
    u64 r0, r1, r2, r3, r4, r5, r6, r7, c;
    u128 s;
    s = (u128)a0 * a1;            r1 = (u64)s; c = (u64)(s >> 64);
    s = (u128)a0 * a2 + c;        r2 = (u64)s; c = (u64)(s >> 64);
    s = (u128)a0 * a3 + c;        r3 = (u64)s; r4 = (u64)(s >> 64);
    s = (u128)a1 * a2 + r3;       r3 = (u64)s; c = (u64)(s >> 64);
    s = (u128)a1 * a3 + r4 + c;   r4 = (u64)s; r5 = (u64)(s >> 64);
    s = (u128)a2 * a3 + r5;       r5 = (u64)s; r6 = (u64)(s >> 64);
    r7 = r6 >> 63;
    r6 = (r6 << 1) | (r5 >> 63); r5 = (r5 << 1) | (r4 >> 63); r4 = (r4 << 1) | (r3 >> 63);
    r3 = (r3 << 1) | (r2 >> 63); r2 = (r2 << 1) | (r1 >> 63); r1 <<= 1;
    s = (u128)a0 * a0;            r0 = (u64)s; c = (u64)(s >> 64);
    s = (u128)r1 + c;             r1 = (u64)s; c = (u64)(s >> 64);
    s = (u128)a1 * a1 + r2 + c;   r2 = (u64)s; c = (u64)(s >> 64);
    s = (u128)r3 + c;             r3 = (u64)s; c = (u64)(s >> 64);
    s = (u128)a2 * a2 + r4 + c;   r4 = (u64)s; c = (u64)(s >> 64);
    s = (u128)r5 + c;             r5 = (u64)s; c = (u64)(s >> 64);
    s = (u128)a3 * a3 + r6 + c;   r6 = (u64)s; c = (u64)(s >> 64);
    r7 += c;
    /* reduction: row i clears limb i; its carry-out joins limb i+4 (running carry `hc` into the next row's top) */
    u64 m, hc = 0;
#define SQ_ROW(x0, x1, x2, x3, x4)                                                     \
    do {                                                                               \
        m = (x0) * ninv;                                                               \
        s = (u128)m * p0 + (x0);          c = (u64)(s >> 64);                          \
        s = (u128)m * p1 + (x1) + c;      (x1) = (u64)s; c = (u64)(s >> 64);           \
        s = (u128)m * p2 + (x2) + c;      (x2) = (u64)s; c = (u64)(s >> 64);           \
        s = (u128)m * p3 + (x3) + c;      (x3) = (u64)s; c = (u64)(s >> 64);           \
        s = (u128)(x4) + c + hc;          (x4) = (u64)s; hc = (u64)(s >> 64);          \
    } while (0)
    SQ_ROW(r0, r1, r2, r3, r4);
    SQ_ROW(r1, r2, r3, r4, r5);
    SQ_ROW(r2, r3, r4, r5, r6);
    SQ_ROW(r3, r4, r5, r6, r7);
#undef SQ_ROW
    /* a < p < 2^255  =>  (a^2 + M p)/R < 2p < 2^256: hc == 0 here */
    u64 d0, d1, d2, d3;
    unsigned long long t;
    unsigned char bw = _subborrow_u64(0, r4, p0, &t); d0 = t;
    bw = _subborrow_u64(bw, r5, p1, &t); d1 = t;
    bw = _subborrow_u64(bw, r6, p2, &t); d2 = t;
    bw = _subborrow_u64(bw, r7, p3, &t); d3 = t;
    r->l[0] = bw ? r4 : d0; r->l[1] = bw ? r5 : d1; r->l[2] = bw ? r6 : d2; r->l[3] = bw ? r7 : d3;
}

/* x^e, left-to-right square-and-multiply over the exponent bits (ark-ff Field::pow, dep). */
static void fe_pow_u64(const field_t *F, fe *r, const fe *x, u64 e) {
    if (e == 0) { *r = F->one; return; }
    int top = 63 - __builtin_clzll(e);
    fe acc = *x;                                  /* leading one bit: 1^2 * x */
    for (int i = top - 1; i >= 0; i--) {
        fe_sqr(F, &acc, &acc);
        if ((e >> i) & 1) fe_mul(F, &acc, &acc, x);
    }
    *r = acc;
}

static void fe_pow4(const field_t *F, fe *r, const fe *x, const u64 e[4]) {
    fe acc = F->one;
    for (int i = 255; i >= 0; i--) {
        fe_mul(F, &acc, &acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) fe_mul(F, &acc, &acc, x);
    }
    *r = acc;
}

static void fe_inv(const field_t *F, fe *r, const fe *x) {
    u64 e[4], two[4] = {2, 0, 0, 0};
    sub4(e, F->p, two);
    fe_pow4(F, r, x, e);
}

int oref_field_init(field_t *F, const u64 p[4]) {
    memcpy(F->p, p, 32);
    F->nocarry = (p[3] >> 63) == 0;
    u64 inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - p[0] * inv;   /* Newton: p^{-1} mod 2^64 */
    F->ninv = (u64)0 - inv;
    /* one = 2^256 mod p by doubling 1 256 times; r2 by 256 more doublings of one. */
    fe x;
    memset(&x, 0, sizeof x);
    x.l[0] = 1;
    for (int i = 0; i < 512; i++) {
        u64 t[4];
        u64 c = add4(t, x.l, x.l);
        if (c || ge4(t, F->p)) sub4(t, t, F->p);
        memcpy(x.l, t, 32);
        if (i == 255) F->one = x;
    }
    F->r2 = x;
    return 0;
}

field_t *oref_field_new(const u64 p[4]) {
    field_t *F = (field_t *)calloc(1, sizeof *F);
    oref_field_init(F, p);
    return F;
}
void oref_field_free(field_t *F) { free(F); }

/* canonical integer (4 LE limbs) <-> Montgomery */
void oref_to_mont(const field_t *F, u64 *out, const u64 *in, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fe a;
        memcpy(a.l, in + 4 * i, 32);
        if (ge4(a.l, F->p)) sub4(a.l, a.l, F->p);      /* inputs < 2^256 < 3p for our fields: loop */
        while (ge4(a.l, F->p)) sub4(a.l, a.l, F->p);
        fe_mul(F, &a, &a, &F->r2);
        memcpy(out + 4 * i, a.l, 32);
    }
}

void oref_from_mont(const field_t *F, u64 *out, const u64 *in, size_t n) {
    fe one_plain = {{1, 0, 0, 0}};
    for (size_t i = 0; i < n; i++) {
        fe a;
        memcpy(a.l, in + 4 * i, 32);
        fe_mul(F, &a, &a, &one_plain);
        memcpy(out + 4 * i, a.l, 32);
    }
}

/* element-wise products / squares / sums, for testing the arithmetic itself against Python integers */
void oref_fe_mul_batch(const field_t *F, u64 *out, const u64 *a, const u64 *b, size_t n) {
    for (size_t i = 0; i < n; i++) fe_mul(F, (fe *)(out + 4 * i), (const fe *)(a + 4 * i), (const fe *)(b + 4 * i));
}
void oref_fe_sqr_batch(const field_t *F, u64 *out, const u64 *a, size_t n) {
    for (size_t i = 0; i < n; i++) fe_sqr(F, (fe *)(out + 4 * i), (const fe *)(a + 4 * i));
}
void oref_fe_addsub_batch(const field_t *F, u64 *sum, u64 *diff, const u64 *a, const u64 *b, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fe_add(F, (fe *)(sum + 4 * i), (const fe *)(a + 4 * i), (const fe *)(b + 4 * i));
        fe_sub(F, (fe *)(diff + 4 * i), (const fe *)(a + 4 * i), (const fe *)(b + 4 * i));
    }
}

/* ---------------------------------------------------------------- parallel-for */

typedef void (*range_fn)(void *ctx, size_t begin, size_t end);

/* Persistent worker pool (the reference's rayon pool is persistent too): workers sleep on a condition variable
 * between jobs; a job is a static split of [0, n) into `threads` chunks, the caller runs chunk 0 itself.  One
 * job at a time (callers serialise on job_mu), one join per call == one rayon barrier per tree level
 * (R/merkle_tree/mod.rs:458,494). */
#define POOL_MAX 1024
static struct {
    pthread_mutex_t mu, job_mu;
    pthread_cond_t work, done;
    pthread_t tid[POOL_MAX];
    int n_workers;                 /* created so far */
    unsigned long gen;             /* job generation */
    range_fn fn; void *ctx; size_t n, chunk; int parts;   /* current job: parts chunks, worker w runs chunk w+1 */
    int pending;
} pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER,
          {0}, 0, 0, NULL, NULL, 0, 0, 0, 0};

static void *pool_main(void *arg) {
    const int w = (int)(size_t)arg;            /* worker index, runs chunk w + 1 */
    unsigned long seen = 0;
    pthread_mutex_lock(&pool.mu);
    for (;;) {
        while (pool.gen == seen) pthread_cond_wait(&pool.work, &pool.mu);
        seen = pool.gen;
        if (w + 1 < pool.parts) {
            range_fn fn = pool.fn; void *ctx = pool.ctx;
            size_t b = (size_t)(w + 1) * pool.chunk, e = b + pool.chunk;
            if (b > pool.n) b = pool.n;
            if (e > pool.n) e = pool.n;
            pthread_mutex_unlock(&pool.mu);
            fn(ctx, b, e);
            pthread_mutex_lock(&pool.mu);
            if (--pool.pending == 0) pthread_cond_signal(&pool.done);
        }
    }
    return NULL;
}

static void parallel_for(size_t n, int threads, range_fn fn, void *ctx) {
    if (threads < 1) threads = 1;
    if (threads > POOL_MAX) threads = POOL_MAX;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    if (threads == 1) { fn(ctx, 0, n); return; }
    pthread_mutex_lock(&pool.job_mu);
    pthread_mutex_lock(&pool.mu);
    while (pool.n_workers < threads - 1) {
        /* a new worker must not mistake the job about to be posted for an old one: it starts with seen = 0 and
         * gen >= 1 only after the post below, so it simply picks this job up */
        if (pthread_create(&pool.tid[pool.n_workers], NULL, pool_main, (void *)(size_t)pool.n_workers)) break;
        pthread_detach(pool.tid[pool.n_workers]);
        pool.n_workers++;
    }
    int parts = pool.n_workers + 1 < threads ? pool.n_workers + 1 : threads;
    size_t chunk = (n + parts - 1) / parts;
    pool.fn = fn; pool.ctx = ctx; pool.n = n; pool.chunk = chunk; pool.parts = parts; pool.pending = parts - 1;
    pool.gen++;
    pthread_cond_broadcast(&pool.work);
    pthread_mutex_unlock(&pool.mu);
    fn(ctx, 0, chunk < n ? chunk : n);
    pthread_mutex_lock(&pool.mu);
    while (pool.pending > 0) pthread_cond_wait(&pool.done, &pool.mu);
    pthread_mutex_unlock(&pool.mu);
    pthread_mutex_unlock(&pool.job_mu);
}

/* ---------------------------------------------------------------- Poseidon */

#define MAX_T 16

typedef struct {
    field_t F;
    int rate, capacity, full_rounds, partial_rounds;
    u64 alpha;
    fe *ark;   /* [(RF+RP)][t] */
    fe *mds;   /* [t][t] */
} poseidon_t;

poseidon_t *oref_poseidon_new(const u64 p[4], int rate, int capacity, int full_rounds,
                              int partial_rounds, u64 alpha, const u64 *ark_mont, const u64 *mds_mont) {
    int t = rate + capacity;
    if (t > MAX_T || t < 2) return NULL;
    poseidon_t *P = (poseidon_t *)calloc(1, sizeof *P);
    oref_field_init(&P->F, p);
    P->rate = rate; P->capacity = capacity;
    P->full_rounds = full_rounds; P->partial_rounds = partial_rounds; P->alpha = alpha;
    size_t na = (size_t)(full_rounds + partial_rounds) * t, nm = (size_t)t * t;
    P->ark = (fe *)malloc(na * sizeof(fe));
    P->mds = (fe *)malloc(nm * sizeof(fe));
    memcpy(P->ark, ark_mont, na * 32);
    memcpy(P->mds, mds_mont, nm * 32);
    return P;
}

void oref_poseidon_free(poseidon_t *P) {
    if (!P) return;
    free(P->ark); free(P->mds); free(P);
}

const field_t *oref_poseidon_field(const poseidon_t *P) { return &P->F; }

/* PoseidonSponge::permute, R/sponge/poseidon/mod.rs:66-121: ark, s-box, dense MDS each round. */
static void permute(const poseidon_t *P, fe *st) {
    const field_t *F = &P->F;
    int t = P->rate + P->capacity, half = P->full_rounds / 2;
    fe nw[MAX_T];
    for (int r = 0; r < P->full_rounds + P->partial_rounds; r++) {
        for (int i = 0; i < t; i++) fe_add(F, &st[i], &st[i], &P->ark[r * t + i]);      /* :79-83 */
        if (r < half || r >= half + P->partial_rounds) {                                   /* :66-77 */
            for (int i = 0; i < t; i++) fe_pow_u64(F, &st[i], &st[i], P->alpha);
        } else {
            fe_pow_u64(F, &st[0], &st[0], P->alpha);
        }
        for (int i = 0; i < t; i++) {                                                      /* :85-96 */
            fe cur;
            memset(&cur, 0, sizeof cur);
            for (int j = 0; j < t; j++) {
                fe term;
                fe_mul(F, &term, &st[j], &P->mds[i * t + j]);
                fe_add(F, &cur, &cur, &term);
            }
            nw[i] = cur;
        }
        memcpy(st, nw, sizeof(fe) * t);
    }
}

void oref_poseidon_permute(const poseidon_t *P, u64 *state_inout) {
    permute(P, (fe *)state_inout);
}

/* crh::poseidon::CRH::evaluate (R/crh/poseidon/mod.rs:30-40): new sponge, absorb the whole
 * input (absorb_internal, sponge/poseidon/mod.rs:124-153), squeeze one element (:323-345). */
static void crh_one(const poseidon_t *P, const fe *in, size_t len, fe *out) {
    const field_t *F = &P->F;
    int t = P->rate + P->capacity;
    fe st[MAX_T];
    memset(st, 0, sizeof(fe) * t);
    size_t pos = 0;
    while (len - pos > (size_t)P->rate) {           /* more than `rate` remain: fill and permute */
        for (int i = 0; i < P->rate; i++) fe_add(F, &st[P->capacity + i], &st[P->capacity + i], &in[pos + i]);
        permute(P, st);
        pos += P->rate;
    }
    for (size_t i = 0; pos + i < len; i++) fe_add(F, &st[P->capacity + i], &st[P->capacity + i], &in[pos + i]);
    permute(P, st);                                 /* squeeze from Absorbing mode permutes once */
    *out = st[P->capacity];
}

typedef struct { const poseidon_t *P; const fe *in; size_t len; size_t stride; fe *out; } crh_job;
static void crh_range(void *c, size_t b, size_t e) {
    crh_job *j = (crh_job *)c;
    for (size_t i = b; i < e; i++) crh_one(j->P, j->in + i * j->stride, j->len, &j->out[i]);
}

/* n independent CRH evaluations, each over `len` elements (inputs contiguous, `len` apart). */
void oref_poseidon_crh_batch(const poseidon_t *P, const u64 *in, size_t len, u64 *out, size_t n, int threads) {
    crh_job j = {P, (const fe *)in, len, len, (fe *)out};
    parallel_for(n, threads, crh_range, &j);
}

/* TwoToOneCRH::compress (R/crh/poseidon/mod.rs:66-79) over n (left,right) pairs. */
void oref_poseidon_compress_batch(const poseidon_t *P, const u64 *pairs, u64 *out, size_t n, int threads) {
    crh_job j = {P, (const fe *)pairs, 2, 2, (fe *)out};
    parallel_for(n, threads, crh_range, &j);
}

/* MerkleTree::new with Poseidon leaf hash + Poseidon two-to-one, identity converter
 * (R/merkle_tree/mod.rs:411-523).  leaf_nodes[n], non_leaf_nodes[n-1] in heap order. */
int oref_poseidon_merkle(const poseidon_t *leafP, const poseidon_t *nodeP, const u64 *leaves, size_t leaf_len,
                         size_t n, u64 *leaf_nodes, u64 *non_leaf_nodes, int threads) {
    if (n < 2 || (n & (n - 1))) return 1;                                     /* :430-433 */
    oref_poseidon_crh_batch(leafP, leaves, leaf_len, leaf_nodes, n, threads);   /* :417-419 */
    fe *nodes = (fe *)non_leaf_nodes;
    size_t start = n / 2 - 1;
    oref_poseidon_compress_batch(nodeP, leaf_nodes, (u64 *)(nodes + start), n / 2, threads);   /* :454-483 */
    while (start > 0) {                                                       /* :486-515 */
        size_t upper = start;
        start = (start - 1) / 2;
        oref_poseidon_compress_batch(nodeP, (const u64 *)(nodes + upper), (u64 *)(nodes + start),
                                     upper - start, threads);
    }
    return 0;
}

/* ---------------------------------------------------------------- twisted Edwards + Pedersen */

typedef struct { fe x, y, t, z; } te_point;   /* extended coordinates, x=X/Z, y=Y/Z, t=XY/Z */

typedef struct {
    field_t F;
    fe a, d;                  /* curve a*x^2 + y^2 = 1 + d*x^2*y^2 */
    int window_size, num_windows;
    size_t n_gens;            /* window_size*num_windows */
    fe *gx, *gy;              /* generators[w][j] flattened, affine */
    size_t n_rand;            /* randomness generators (0 for plain CRH) */
    fe *rx, *ry;
} pedersen_t;

pedersen_t *oref_pedersen_new(const u64 p[4], const u64 *a_mont, const u64 *d_mont, int window_size,
                              int num_windows, const u64 *gens_xy_mont, size_t n_rand, const u64 *rand_xy_mont) {
    pedersen_t *P = (pedersen_t *)calloc(1, sizeof *P);
    oref_field_init(&P->F, p);
    memcpy(P->a.l, a_mont, 32);
    memcpy(P->d.l, d_mont, 32);
    P->window_size = window_size; P->num_windows = num_windows;
    P->n_gens = (size_t)window_size * num_windows;
    P->gx = (fe *)malloc(P->n_gens * sizeof(fe)); P->gy = (fe *)malloc(P->n_gens * sizeof(fe));
    for (size_t i = 0; i < P->n_gens; i++) {
        memcpy(P->gx[i].l, gens_xy_mont + 8 * i, 32);
        memcpy(P->gy[i].l, gens_xy_mont + 8 * i + 4, 32);
    }
    P->n_rand = n_rand;
    if (n_rand) {
        P->rx = (fe *)malloc(n_rand * sizeof(fe)); P->ry = (fe *)malloc(n_rand * sizeof(fe));
        for (size_t i = 0; i < n_rand; i++) {
            memcpy(P->rx[i].l, rand_xy_mont + 8 * i, 32);
            memcpy(P->ry[i].l, rand_xy_mont + 8 * i + 4, 32);
        }
    }
    return P;
}

void oref_pedersen_free(pedersen_t *P) {
    if (!P) return;
    free(P->gx); free(P->gy); free(P->rx); free(P->ry); free(P);
}

/* Unified twisted-Edwards addition for general a (Hisil-Wong-Carter-Dawson, "add-2008-hwcd"):
 * complete when a is a square and d a non-square.  acc += (x2,y2) affine. */
static void te_add_affine(const pedersen_t *P, te_point *acc, const fe *x2, const fe *y2) {
    const field_t *F = &P->F;
    fe A, B, C, D, E, Fv, G, H, t2, tmp, tmp2;
    fe_mul(F, &A, &acc->x, x2);
    fe_mul(F, &B, &acc->y, y2);
    fe_mul(F, &t2, x2, y2);
    fe_mul(F, &C, &acc->t, &t2);
    fe_mul(F, &C, &C, &P->d);
    D = acc->z;
    fe_add(F, &tmp, &acc->x, &acc->y);
    fe_add(F, &tmp2, x2, y2);
    fe_mul(F, &E, &tmp, &tmp2);
    fe_sub(F, &E, &E, &A);
    fe_sub(F, &E, &E, &B);
    fe_sub(F, &Fv, &D, &C);
    fe_add(F, &G, &D, &C);
    fe_mul(F, &tmp, &P->a, &A);
    fe_sub(F, &H, &B, &tmp);
    fe_mul(F, &acc->x, &E, &Fv);
    fe_mul(F, &acc->y, &G, &H);
    fe_mul(F, &acc->t, &E, &H);
    fe_mul(F, &acc->z, &Fv, &G);
}

static void te_identity(const pedersen_t *P, te_point *r) {
    memset(r, 0, sizeof *r);
    r->y = P->F.one;
    r->z = P->F.one;
}

static void te_to_affine(const pedersen_t *P, const te_point *a, fe *x, fe *y) {
    fe zi;
    fe_inv(&P->F, &zi, &a->z);
    fe_mul(&P->F, x, &a->x, &zi);
    fe_mul(&P->F, y, &a->y, &zi);
}

/* pedersen::CRH::evaluate (R/crh/pedersen/mod.rs:76-129) on an already padded input of `len`
 * bytes; optionally followed by the commitment's randomness sum
 * (R/commitment/pedersen/mod.rs:93-100).  Returns 0, or 1 for "incorrect input length". */
static int pedersen_one(const pedersen_t *P, const uint8_t *in, size_t len, const uint8_t *rand32, fe *ox, fe *oy) {
    size_t nbits = P->n_gens;
    if (len * 8 > nbits) return 1;                                  /* :82-89 (panic) */
    size_t padded = len;
    if (len * 8 < nbits) padded = nbits / 8 > len ? nbits / 8 : len;   /* :94-99 */
    te_point acc;
    te_identity(P, &acc);
    size_t total_bits = padded * 8;
    for (size_t w = 0; w < (size_t)P->num_windows; w++) {           /* chunks(WINDOW_SIZE).zip(generators) :113-124 */
        for (size_t j = 0; j < (size_t)P->window_size; j++) {
            size_t k = w * P->window_size + j;
            if (k >= total_bits) break;
            size_t byte = k >> 3;
            int bit = byte < len ? (in[byte] >> (k & 7)) & 1 : 0;   /* bytes_to_bits :200-209 */
            if (bit) te_add_affine(P, &acc, &P->gx[k], &P->gy[k]);
        }
    }
    if (rand32) {
        for (size_t k = 0; k < P->n_rand && k < 256; k++)
            if ((rand32[k >> 3] >> (k & 7)) & 1) te_add_affine(P, &acc, &P->rx[k], &P->ry[k]);
    }
    te_to_affine(P, &acc, ox, oy);                                  /* :128 into() */
    return 0;
}

typedef struct {
    const pedersen_t *P; const uint8_t *in; size_t len, stride; const uint8_t *rand; u64 *out; int err;
} ped_job;
static void ped_range(void *c, size_t b, size_t e) {
    ped_job *j = (ped_job *)c;
    for (size_t i = b; i < e; i++) {
        fe x, y;
        if (pedersen_one(j->P, j->in + i * j->stride, j->len, j->rand ? j->rand + 32 * i : NULL, &x, &y)) {
            j->err = 1;
            return;
        }
        memcpy(j->out + 8 * i, x.l, 32);
        memcpy(j->out + 8 * i + 4, y.l, 32);
    }
}

/* n hashes of `len` bytes each (`stride` apart); out = n x (x,y) Montgomery. rand32: n x 32-byte
 * little-endian canonical scalars for commit, or NULL for the CRH. */
int oref_pedersen_batch(const pedersen_t *P, const uint8_t *in, size_t len, size_t stride, const uint8_t *rand32,
                        u64 *out_xy, size_t n, int threads) {
    ped_job j = {P, in, len, stride, rand32, out_xy, 0};
    parallel_for(n, threads, ped_range, &j);
    return j.err;
}

/* serialize_uncompressed of an affine TE point: canonical x || y, 32-byte LE each (dep). */
static void point_bytes(const field_t *F, const u64 *xy_mont, uint8_t out[64]) {
    u64 plain[8];
    oref_from_mont(F, plain, xy_mont, 2);
    memcpy(out, plain, 64);       /* little-endian host */
}

typedef struct { const pedersen_t *P; const u64 *children; u64 *out; int err; } pnode_job;
static void pnode_range(void *c, size_t b, size_t e) {
    pnode_job *j = (pnode_job *)c;
    size_t half_bytes = j->P->n_gens / 2 / 8;
    size_t buf_len = 2 * half_bytes;
    uint8_t *buf = (uint8_t *)malloc(buf_len + 128);
    for (size_t i = b; i < e; i++) {
        uint8_t lr[128];
        point_bytes(&j->P->F, j->children + 16 * i, lr);
        point_bytes(&j->P->F, j->children + 16 * i + 8, lr + 64);
        memset(buf, 0, buf_len);                                       /* R/crh/pedersen/mod.rs:173-180 */
        memcpy(buf, lr, buf_len < 128 ? buf_len : 128);
        fe x, y;
        if (pedersen_one(j->P, buf, buf_len, NULL, &x, &y)) { j->err = 1; break; }
        memcpy(j->out + 8 * i, x.l, 32);
        memcpy(j->out + 8 * i + 4, y.l, 32);
    }
    free(buf);
}

/* TwoToOneCRH::compress over n (left,right) affine point pairs (R/crh/pedersen/mod.rs:187-197). */
int oref_pedersen_compress_batch(const pedersen_t *P, const u64 *children_xy, u64 *out_xy, size_t n, int threads) {
    pnode_job j = {P, children_xy, out_xy, 0};
    parallel_for(n, threads, pnode_range, &j);
    return j.err;
}

/* Byte-leaf Pedersen tree: LeafHash = pedersen::CRH, ByteDigestConverter, TwoToOneHash =
 * pedersen::TwoToOneCRH (R/merkle_tree/tests/mod.rs:19-33).  Digests are (x,y) Montgomery. */
int oref_pedersen_merkle(const pedersen_t *leafP, const pedersen_t *nodeP, const uint8_t *leaves, size_t leaf_len,
                         size_t n, u64 *leaf_nodes_xy, u64 *non_leaf_nodes_xy, int threads) {
    if (n < 2 || (n & (n - 1))) return 1;
    if (oref_pedersen_batch(leafP, leaves, leaf_len, leaf_len, NULL, leaf_nodes_xy, n, threads)) return 2;
    size_t start = n / 2 - 1;
    if (oref_pedersen_compress_batch(nodeP, leaf_nodes_xy, non_leaf_nodes_xy + 8 * start, n / 2, threads)) return 2;
    while (start > 0) {
        size_t upper = start;
        start = (start - 1) / 2;
        if (oref_pedersen_compress_batch(nodeP, non_leaf_nodes_xy + 8 * upper, non_leaf_nodes_xy + 8 * start,
                                         upper - start, threads)) return 2;
    }
    return 0;
}

/* Mixed tree (BASELINE config 5): LeafHash = PedersenCRHCompressor<_, TECompressor, W>
 * (x-coordinate, R/crh/injective_map/mod.rs:22-62), identity converter, Poseidon two-to-one. */
int oref_mixed_merkle(const pedersen_t *leafP, const poseidon_t *nodeP, const uint8_t *leaves, size_t leaf_len,
                      size_t n, u64 *leaf_nodes, u64 *non_leaf_nodes, int threads) {
    if (n < 2 || (n & (n - 1))) return 1;
    u64 *xy = (u64 *)malloc(n * 64);
    if (oref_pedersen_batch(leafP, leaves, leaf_len, leaf_len, NULL, xy, n, threads)) { free(xy); return 2; }
    for (size_t i = 0; i < n; i++) memcpy(leaf_nodes + 4 * i, xy + 8 * i, 32);
    free(xy);
    fe *nodes = (fe *)non_leaf_nodes;
    size_t start = n / 2 - 1;
    oref_poseidon_compress_batch(nodeP, leaf_nodes, (u64 *)(nodes + start), n / 2, threads);
    while (start > 0) {
        size_t upper = start;
        start = (start - 1) / 2;
        oref_poseidon_compress_batch(nodeP, (const u64 *)(nodes + upper), (u64 *)(nodes + start), upper - start, threads);
    }
    return 0;
}

/* ---------------------------------------------------------------- Bowe-Hopwood Pedersen CRH */

/* bowe_hopwood::CRH::evaluate (R/crh/bowe_hopwood/mod.rs:115-185) over the generators of a pedersen_t
 * (generators[w][j] flattened; window_size = chunks per segment).  Output: x-coordinate (Montgomery).
 * Returns 1 for "incorrect input bitlength" (:121-129). */
static int bh_one(const pedersen_t *P, const uint8_t *in, size_t len, fe *ox) {
    if (len * 8 > P->n_gens * 3) return 1;
    size_t nbits = len * 8, nchunks = (nbits + 2) / 3;                  /* padded to a multiple of 3 with zeros (:133-140) */
    te_point acc;
    te_identity(P, &acc);
    for (size_t k = 0; k < nchunks && k < P->n_gens; k++) {             /* chunk k <-> generators[k / WS][k % WS] */
        int c[3];
        for (int b = 0; b < 3; b++) {
            size_t bit = 3 * k + b;
            c[b] = bit < nbits ? (in[bit >> 3] >> (bit & 7)) & 1 : 0;
        }
        /* encoded = g (+ g if c0) (+ 2g if c1), negated if c2 (:165-177) */
        te_point enc;
        te_identity(P, &enc);
        int mult = 1 + c[0] + 2 * c[1];
        for (int r = 0; r < mult; r++) te_add_affine(P, &enc, &P->gx[k], &P->gy[k]);
        fe ex, ey;
        te_to_affine(P, &enc, &ex, &ey);
        if (c[2]) { fe z; memset(&z, 0, sizeof z); fe_sub(&P->F, &ex, &z, &ex); }
        te_add_affine(P, &acc, &ex, &ey);
    }
    fe y;
    te_to_affine(P, &acc, ox, &y);
    return 0;
}

typedef struct { const pedersen_t *P; const uint8_t *in; size_t len, stride; u64 *out; int err; } bh_job;
static void bh_range(void *c, size_t b, size_t e) {
    bh_job *j = (bh_job *)c;
    for (size_t i = b; i < e; i++) {
        fe x;
        if (bh_one(j->P, j->in + i * j->stride, j->len, &x)) { j->err = 1; return; }
        memcpy(j->out + 4 * i, x.l, 32);
    }
}
int oref_bowe_hopwood_batch(const pedersen_t *P, const uint8_t *in, size_t len, size_t stride, u64 *out_x, size_t n, int threads) {
    bh_job j = {P, in, len, stride, out_x, 0};
    parallel_for(n, threads, bh_range, &j);
    return j.err;
}
