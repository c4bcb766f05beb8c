/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under fast_monte_carlo_b200/ may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs load it (through oracle/c_oracle.py).
 *
 * Plain-C, scalar, one-game-at-a-time restatement of the reference's play-by-play engine
 * (/root/reference/fast_monte_carlo_cfb.py, "FMC") for the hot path of SURVEY.md section 8:
 *
 *   tree arithmetic   XGBoost 3.0.4 / scikit-learn 1.5.2 predictors behind FMC:745, 757, 784-809
 *                     (un-vendored third-party code: restated from the published algorithms,
 *                      SURVEY Appendix D; unpruned forests, generic row view)
 *   feature row       _fill_row                       FMC:996-1021
 *   play call         pass_prob_v1                    FMC:719-735  (play_model.xgb policy optional)
 *   pass outcome      pass_stage1_proba/stage2_proba  FMC:739-770
 *   yardage           sample_{pass,rush}_yards, sample_sack_loss   FMC:817-852
 *                     QuantileYards.sample            sim_helpers.py:32-38 (optional sampler)
 *   modifiers         matchup_bias .. explosive_prob  FMC:431-472
 *   special teams     field_goal_prob, attempt_fg, attempt_punt    FMC:858-896
 *   state machine     advance_down, change_possession, tick_clock  FMC:927-968
 *   one play          simulate_play                   FMC:1026-1257
 *   fourth down       go_for_it_prob, handle_fourth   FMC:1336-1421
 *   one game          simulate_game                   FMC:1428-1464
 *   players           sample_qb/rusher/target FMC:625-635, focus tracking + per-player box FMC:1058-1075,
 *                     1105-1148, 1163-1192, 1203-1249 (fo_simulate_players; usage tables are inputs)
 *
 * Pinning: tests/test_oracle_golden.py replays tests/golden/ref_trajectories.npz, which was
 * produced by the UNMODIFIED reference module run in the build container
 * (oracle/ref_harness.py + tests/golden/make_golden.py), and checks every per-iteration state.
 * The scikit-learn quantile models are pinned bit-for-bit against the live pipelines
 * (tests/golden/sklearn_quantiles.npz).  The XGBoost boosters are "parity unpinned": xgboost is
 * not installed, so their evaluation here (and in oracle/fake_xgboost.py, which the golden run
 * used) is a restatement only.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp; no fast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FO_N_MODELS 7
#define FO_PASS_STAGE1 0
#define FO_PASS_STAGE2 1
#define FO_PASS_YARDS 2
#define FO_RUN_YARDS 3
#define FO_SACK_YARDS 4
#define FO_PLAY_MODEL 5
#define FO_RUN_FUMBLE 6

#define FO_KIND_XGB 0
#define FO_KIND_SKL 1

#define FO_N_SLOTS 16
#define FO_MAX_ITERS 360
#define FO_TRACE_COLS 8

/* injected-stream slot layout (oracle/ref_harness.py SLOT) */
enum { S_U_CALL = 0, S_U_COMP, S_Z_YARDS, S_U_EX, S_U_BOOST, S_U_FIN, S_U_S2, S_Z_INT,
       S_U_GO, S_U_FG, S_Z_GROSS, S_Z_RET, S_U_TB, S_U_P1, S_U_WR, S_U_YQ };

typedef struct {
    int loaded, kind, n_outputs, n_features, num_base, n_num, zero_is_missing;
    double base[8];
    double scale;
    int n_nodes, n_trees;
    int *feat, *left, *right, *root, *out;
    float *thr;
    unsigned char *dl;
    double *value;
    /* play_model only */
    int n_scaled;
    int scaler_cols[16];
    double scaler_mean[16], scaler_scale[16];
} FoForest;

static FoForest g_forest[FO_N_MODELS];

typedef struct {
    double sp[2][3];             /* [team][RATING, OFFENSE, DEFENSE]  (FMC:1625-1644) */
    int active[FO_N_MODELS][2];  /* hot one-hot columns for the single "Unknown" player, -1 = none */
    int coach_col[2];            /* play_model coach column per team, -1 = none */
    int policy;                  /* 0 = pass_prob_v1 (FMC:719), 1 = play_model.xgb softmax[pass] */
    double play_temp;            /* _PLAY_TEMP (FMC:50, 335-337) */
    int sampler;                 /* 0 = FMC Normal sampler, 1 = sim_helpers.QuantileYards.sample */
    double qy_noise;             /* sim_helpers noise (default 0.5) */
    int stage2_mode;             /* 0 = stand-in probabilities, 1 = booster FO_PASS_STAGE2 */
    double standin[3];           /* raw [incomplete, intercepted, sack] as float32-exact doubles */
    int pass_class;              /* policy 1: index of "pass" among the play model's classes: 1 for play_model.xgb
                                    ([field_goal, pass, punt, run, timeout]), label_encoder.pkl order for
                                    play_model.json ([pass, run] -> 0; FMC:333, 423) */
} FoConfig;

/* Usage tables of one team (TeamContext.qb_share / rush_share / target_share, FMC:262-264) reduced to what
 * the engine reads: the shares Generator.choice gets (FMC:627, 631, 635), whether a name is in the team's
 * focus track set (FMC:1062-1063, 1204) and which one-hot column the name lights in every model. */
#define FO_MAX_USAGE 32
#define FO_PLAYER_FIELDS 6   /* yds, att|tgt, comp|rec, td, INT, sacks */
typedef struct {
    int n;
    double share[FO_MAX_USAGE];
    int slot[FO_MAX_USAGE];               /* output slot of a tracked name, -1 = not tracked */
    int col[FO_N_MODELS][FO_MAX_USAGE];   /* hot column of the name in model m (-1 = not a category) */
} FoUsage;
typedef struct { FoUsage role[3]; } FoTeamUsage;   /* 0 passer, 1 rusher, 2 target */

static void *dup_mem(const void *p, size_t n) {
    void *q = malloc(n ? n : 1);
    if (q && n) memcpy(q, p, n);
    return q;
}

int fo_load_forest(int id, int kind, int n_outputs, int n_features, int num_base, int n_num,
                   int zero_is_missing, const double *base, double scale, int n_nodes,
                   const int *feat, const float *thr, const int *left, const int *right,
                   const unsigned char *dl, const double *value, int n_trees, const int *root,
                   const int *out) {
    if (id < 0 || id >= FO_N_MODELS || n_outputs > 8) return -1;
    FoForest *f = &g_forest[id];
    if (f->loaded) {
        free(f->feat); free(f->left); free(f->right); free(f->root); free(f->out);
        free(f->thr); free(f->dl); free(f->value);
    }
    memset(f, 0, sizeof(*f));
    f->kind = kind; f->n_outputs = n_outputs; f->n_features = n_features; f->num_base = num_base;
    f->n_num = n_num; f->zero_is_missing = zero_is_missing; f->scale = scale;
    for (int k = 0; k < n_outputs; ++k) f->base[k] = base[k];
    f->n_nodes = n_nodes; f->n_trees = n_trees;
    f->feat = dup_mem(feat, sizeof(int) * n_nodes);
    f->left = dup_mem(left, sizeof(int) * n_nodes);
    f->right = dup_mem(right, sizeof(int) * n_nodes);
    f->thr = dup_mem(thr, sizeof(float) * n_nodes);
    f->dl = dup_mem(dl, n_nodes);
    f->value = dup_mem(value, sizeof(double) * n_nodes);
    f->root = dup_mem(root, sizeof(int) * n_trees);
    f->out = dup_mem(out, sizeof(int) * n_trees);
    f->loaded = 1;
    return 0;
}

int fo_set_scaler(int id, int n, const int *cols, const double *mean, const double *scale) {
    if (id < 0 || id >= FO_N_MODELS || n > 16) return -1;
    FoForest *f = &g_forest[id];
    f->n_scaled = n;
    for (int i = 0; i < n; ++i) { f->scaler_cols[i] = cols[i]; f->scaler_mean[i] = mean[i]; f->scaler_scale[i] = scale[i]; }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* tree arithmetic                                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double num[17];
    int active[2];
} FoRow;

static inline float row_value(const FoForest *f, int col, const FoRow *r) {
    if (col >= f->num_base && col < f->num_base + f->n_num) return (float)r->num[col - f->num_base];
    return (col == r->active[0] || col == r->active[1]) ? 1.0f : 0.0f;
}

static inline double walk(const FoForest *f, int t, const FoRow *r) {
    int i = f->root[t];
    while (f->left[i] >= 0) {
        float v = row_value(f, f->feat[i], r);
        int go_left;
        if (f->kind == FO_KIND_XGB) {
            if (f->zero_is_missing && v == 0.0f) go_left = f->dl[i] != 0;  /* absent from the CSR row */
            else if (v != v) go_left = f->dl[i] != 0;                      /* NaN: xgboost's missing value */
            else go_left = v < f->thr[i];
        } else {
            go_left = v <= f->thr[i];
        }
        i = go_left ? f->left[i] : f->right[i];
    }
    return f->value[i];
}

/* raw margins in the model's own precision / order; out[n_outputs] */
static void margins(const FoForest *f, const FoRow *r, int tree_begin, int tree_end, double *out) {
    if (tree_end < 0 || tree_end > f->n_trees) tree_end = f->n_trees;
    if (f->kind == FO_KIND_XGB) {
        float acc[8];
        for (int k = 0; k < f->n_outputs; ++k) acc[k] = (float)f->base[k];
        for (int t = tree_begin; t < tree_end; ++t) acc[f->out[t]] += (float)walk(f, t, r);
        for (int k = 0; k < f->n_outputs; ++k) out[k] = (double)acc[k];
    } else {
        double acc[8];
        for (int k = 0; k < f->n_outputs; ++k) acc[k] = f->base[k];
        for (int t = tree_begin; t < tree_end; ++t) {
            double term = f->scale * walk(f, t, r);
            acc[f->out[t]] += term;
        }
        for (int k = 0; k < f->n_outputs; ++k) out[k] = acc[k];
    }
}

/* correctly-rounded-by-construction float exp: double exp, one rounding to float */
static inline float expf_cr(float x) { return (float)exp((double)x); }

static inline float sigmoid_f32(float m) { return 1.0f / (expf_cr(-m) + 1.0f); }

static void softmax_f32(const float *m, int n, float *p) {
    float wmax = m[0];
    for (int i = 1; i < n; ++i) wmax = fmaxf(m[i], wmax);
    double wsum = 0.0;
    for (int i = 0; i < n; ++i) { p[i] = expf_cr(m[i] - wmax); wsum += p[i]; }
    for (int i = 0; i < n; ++i) p[i] /= (float)wsum;
}

int fo_predict(int id, long n, const double *num, const int *active, int tree_begin, int tree_end,
               double *out) {
    if (id < 0 || id >= FO_N_MODELS || !g_forest[id].loaded) return -1;
    const FoForest *f = &g_forest[id];
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        FoRow r;
        memset(&r, 0, sizeof(r));
        for (int k = 0; k < f->n_num; ++k) r.num[k] = num[i * 17 + k];
        r.active[0] = active[i * 2]; r.active[1] = active[i * 2 + 1];
        margins(f, &r, tree_begin, tree_end, out + i * f->n_outputs);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* counter-based RNG: Philox4x32-10 (Salmon et al. 2011) + AS241 inverse normal                 */
/* ------------------------------------------------------------------------------------------ */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void fo_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { philox4x32_10(ctr, key, out); }

static inline double u01(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }

/* Wichura (1988) algorithm AS241, PPND16 */
static double ppnd16(double p) {
    static const double a[8] = {3.3871328727963666080e0, 1.3314166789178437745e+2, 1.9715909503065514427e+3,
        1.3731693765509461125e+4, 4.5921953931549871457e+4, 6.7265770927008700853e+4,
        3.3430575583588128105e+4, 2.5090809287301226727e+3};
    static const double b[8] = {1.0, 4.2313330701600911252e+1, 6.8718700749205790830e+2, 5.3941960214247511077e+3,
        2.1213794301586595867e+4, 3.9307895800092710610e+4, 2.8729085735721942674e+4,
        5.2264952788528545610e+3};
    static const double c[8] = {1.42343711074968357734e0, 4.63033784615654529590e0, 5.76949722146069140550e0,
        3.64784832476320460504e0, 1.27045825245236838258e0, 2.41780725177450611770e-1,
        2.27238449892691845833e-2, 7.74545014278341407640e-4};
    static const double d[8] = {1.0, 2.05319162663775882187e0, 1.67638483018380384940e0, 6.89767334985100004550e-1,
        1.48103976427480074590e-1, 1.51986665636164571966e-2, 5.47593808499534494600e-4,
        1.05075007164441684324e-9};
    static const double e[8] = {6.65790464350110377720e0, 5.46378491116411436990e0, 1.78482653991729133580e0,
        2.96560571828504891230e-1, 2.65321895265761230930e-2, 1.24266094738807843860e-3,
        2.71155556874348757815e-5, 2.01033439929228813265e-7};
    static const double f[8] = {1.0, 5.99832206555887937690e-1, 1.36929880922735805310e-1, 1.48753612908506148525e-2,
        7.86869131145613259100e-4, 1.84631831751005468180e-5, 1.42151175831644588870e-7,
        2.04426310338993978564e-15};
    double q = p - 0.5, r, num, den;
    const double *n_, *d_;
    if (fabs(q) <= 0.425) {
        r = 0.180625 - q * q;
        num = a[7]; den = b[7];
        for (int i = 6; i >= 0; --i) { num = num * r + a[i]; den = den * r + b[i]; }
        return q * num / den;
    }
    r = q < 0.0 ? p : 1.0 - p;
    r = sqrt(-log(r));
    if (r <= 5.0) { r -= 1.6; n_ = c; d_ = d; } else { r -= 5.0; n_ = e; d_ = f; }
    num = n_[7]; den = d_[7];
    for (int i = 6; i >= 0; --i) { num = num * r + n_[i]; den = den * r + d_[i]; }
    double v = num / den;
    return q < 0.0 ? -v : v;
}

double fo_ppnd16(double p) { return ppnd16(p); }

typedef struct {
    int mode;                 /* 0 injected, 1 philox */
    const double *rec;        /* injected: [FO_MAX_ITERS][16] of this game */
    uint32_t key[2];
    uint32_t ctr_game[3];     /* game_lo, game_hi, matchup */
    int iter;
    uint32_t words[16];
    int have[4];
} FoRng;

static double draw(FoRng *g, int slot, int normal) {
    if (g->mode == 0) return g->rec[(size_t)g->iter * FO_N_SLOTS + slot];
    int blk = slot >> 2;
    if (!g->have[blk]) {
        uint32_t ctr[4] = {g->ctr_game[0], g->ctr_game[1], g->ctr_game[2], ((uint32_t)g->iter << 2) | (uint32_t)blk};
        philox4x32_10(ctr, g->key, g->words + 4 * blk);
        g->have[blk] = 1;
    }
    double u = u01(g->words[slot]);
    return normal ? ppnd16(u) : u;
}

/* ------------------------------------------------------------------------------------------ */
/* game engine                                                                                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int offense;              /* team index 0/1 currently on offense (GameState.offense) */
    int sec, down;
    double dist, ytg;
    int period, going;
    int score[2];
} FoState;

typedef struct {
    const FoConfig *cfg;
    /* per-orientation constants, index = offense team */
    double bias[2], ymul[2], mz[2], tanh35[2];
    long plays, iters;
    long n_pass, n_comp, n_inc, n_int, n_sack, n_run, n_td, n_fga, n_fg, n_punt, n_go;
    /* players mode (fo_simulate_players) */
    const FoTeamUsage *usage;     /* [2] or NULL */
    double cdf[2][3][FO_MAX_USAGE];
    int n_slots;
    double *box;                  /* this game's [2][n_slots][FO_PLAYER_FIELDS] or NULL */
    int cur[3];                   /* usage entries sampled for the current play: passer, rusher, target */
} FoGame;

/* FMC:97 */
static inline double softclip(double x, double lo, double hi) {
    double m = (x < hi) ? x : hi;    /* min(hi, x) */
    return (m > lo) ? m : lo;        /* max(lo, m) */
}
static inline double pymax(double a, double b) { return (b > a) ? b : a; }   /* max(a, b) */
static inline double pymin(double a, double b) { return (b < a) ? b : a; }   /* min(a, b) */

/* FMC:943-953 */
static void change_possession(FoState *s, int has_spot, double spot) {
    s->offense ^= 1;
    s->down = 1;
    s->dist = 10.0;
    s->going = 0;
    s->ytg = has_spot ? spot : 100.0 - s->ytg;
}

/* FMC:932-941 (first_down_reset FMC:927-930 is called with gained=0) */
static void advance_down(FoState *s, double gained) {
    s->ytg = pymax(0.0, s->ytg - gained);
    if (gained + 1e-6 >= s->dist) {
        s->down = 1;
        s->dist = 10.0;
        s->ytg = pymax(0.0, s->ytg - 0);
    } else {
        s->down += 1;
        s->dist -= gained;
        if (s->down > 4) change_possession(s, 0, 0.0);
    }
}

/* FMC:956-968 */
static void tick_clock(FoState *s, int base) {
    int v = s->sec - base;
    s->sec = v > 0 ? v : 0;
    int old = s->period;
    s->period = s->sec > 0 ? 4 - ((s->sec - 1) / 900) : 4;
    if (s->period != old && s->period == 3) change_possession(s, 1, 75.0);
}

/* FMC:719-735 */
static double pass_prob_v1(int down, double distance, double ytg, int sec, int sd) {
    double base = 0.53;
    if (down == 1) base += 0.02 + 0.010 * pymax(0, distance - 10) / 10;
    if (down == 2) base += 0.12 + 0.020 * pymax(0, distance - 7) / 10;
    if (down == 3) base += 0.28 + 0.030 * pymax(0, distance - 5) / 10;
    if (down == 4) base += 0.45 + 0.035 * pymax(0, distance - 3) / 10;
    if (ytg <= 10) base -= 0.05;
    if (ytg <= 5) base -= 0.03;
    int two_min = (sec % 1800) <= 120;
    if (two_min && sd < 0) base += 0.22;
    if (sec < 600 && sd < 0) base += 0.06;
    return softclip(base, 0.10, 0.95);
}

/* FMC:467-472 */
static double explosive_prob(double mz, double ytg) {
    double base = 0.03 + 0.05 * mz;
    if (ytg > 60) base += 0.02;
    if (ytg > 40) base += 0.01;
    return softclip(base, 0.01, 0.12);
}
/* FMC:444-457 */
static double rz_finish_prob_pass(double ytg, double tanh35, int down) {
    double base = 0.32 + 0.30 * (pymax(0.0, 7.0 - ytg) / 7.0);
    int dl = 4 - down; if (dl < 0) dl = 0;
    base += 0.03 * dl;
    double tilt = 0.08 * tanh35;
    return softclip(base + tilt, 0.22, 0.68);
}
static double rz_finish_prob_run(double ytg, double tanh35, int down) {
    double base = 0.30 + 0.30 * (pymax(0.0, 7.0 - ytg) / 7.0);
    int dl = 4 - down; if (dl < 0) dl = 0;
    base += 0.04 * dl;
    double tilt = 0.07 * tanh35;
    return softclip(base + tilt, 0.20, 0.62);
}

/* FMC:858-865 */
static double field_goal_prob(double d) {
    if (d < 30) return 0.96;
    if (d < 40) return 0.92;
    if (d < 50) return 0.78;
    if (d <= 55) return 0.50;
    return 0.25;
}

/* FMC:1336-1378 */
static double go_for_it_prob(double ytg, double dist, int sd, int sec) {
    if (sec < 300 && sd < 0) return (ytg > 38) ? 0.90 : 0.75;
    double p = 0.0;
    if (ytg > 80) { if (dist <= 1) p = 0.15; else if (dist <= 2) p = 0.05; }
    else if (ytg > 65) { if (dist <= 1) p = 0.30; else if (dist <= 2) p = 0.15; }
    else if (ytg > 50) { if (dist <= 1) p = 0.60; else if (dist <= 2) p = 0.40; else if (dist <= 3) p = 0.20; }
    else if (ytg > 35) { if (dist <= 1) p = 0.85; else if (dist <= 2) p = 0.65; else if (dist <= 3) p = 0.40; else if (dist <= 4) p = 0.25; }
    else if (ytg > 20) { if (dist <= 1) p = 0.75; else if (dist <= 2) p = 0.50; else if (dist <= 3) p = 0.30; }
    else if (ytg > 10) { if (dist <= 1) p = 0.70; else if (dist <= 2) p = 0.45; }
    else { if (dist <= 2) p = 0.85; else if (dist <= 4) p = 0.40; }
    if (sec < 300 && sd > 0) p *= 0.85;
    return softclip(p, 0.0, 1.0);
}

/* FMC:996-1021 -- the 17 numerics of the reusable row */
static void fill_row(FoRow *r, const FoGame *G, const FoState *s, int sd, int model) {
    const FoConfig *c = G->cfg;
    int off = s->offense, de = off ^ 1;
    r->num[0] = (double)s->down;
    r->num[1] = s->dist;
    r->num[2] = s->ytg;
    r->num[3] = (s->ytg <= 20) ? 1.0 : 0.0;
    r->num[4] = (double)sd;
    r->num[5] = (double)s->sec;
    r->num[6] = 3.0;                       /* timeouts are never spent (FMC:911-912) */
    r->num[7] = 3.0;
    r->num[8] = c->sp[off][0];
    r->num[9] = c->sp[off][1];
    r->num[10] = c->sp[de][2];
    r->num[11] = c->sp[de][0];
    r->num[12] = (s->dist >= (s->ytg - 0.5)) ? 1.0 : 0.0;
    r->num[13] = (s->down == 4 && s->dist <= 2.0) ? 1.0 : 0.0;
    r->num[14] = (s->ytg <= 33) ? 1.0 : 0.0;
    r->num[15] = (s->sec > 1800) ? 1.0 : 2.0;
    r->num[16] = ((s->sec % 1800) <= 120) ? 1.0 : 0.0;
    r->active[0] = c->active[model][0];
    r->active[1] = c->active[model][1];
    if (G->usage) {      /* passer_name / target_name / rusher_name of the row (FMC:1079-1081, 1216) */
        const FoTeamUsage *u = &G->usage[off];
        if (model == FO_RUN_YARDS) {
            r->active[0] = u->role[1].col[model][G->cur[1]];
            r->active[1] = -1;
        } else {
            r->active[0] = u->role[0].col[model][G->cur[0]];
            r->active[1] = u->role[2].col[model][G->cur[2]];
        }
    }
}

/* Generator.choice(n, p=shares): cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, 'right') */
static int sample_usage(const FoGame *G, int team, int role, double u) {
    const int n = G->usage[team].role[role].n;
    int idx = 0;
    for (int i = 0; i < n; ++i) idx += (G->cdf[team][role][i] <= u);
    return idx < n ? idx : n - 1;
}

/* pstats[team][role][name][field] += v for a tracked name (FMC:1073-1075, 1108-1148, 1163-1192, 1207-1249) */
static void credit(FoGame *G, int team, int role, int field, double v) {
    if (!G->box) return;
    const int slot = G->usage[team].role[role].slot[G->cur[role]];
    if (slot < 0) return;
    G->box[((size_t)team * G->n_slots + slot) * FO_PLAYER_FIELDS + field] += v;
}

/* three quantiles of one family (FMC:780-812) */
static void quants(const FoGame *G, const FoState *s, int sd, int model, double q[3]) {
    FoRow r;
    fill_row(&r, G, s, sd, model);
    margins(&g_forest[model], &r, 0, -1, q);
}

/* yardage samplers: FMC:817-852, or sim_helpers.py:32-38 when cfg->sampler == 1 */
static double sample_yards(const FoGame *G, FoRng *rng, const double q[3], double sig_floor, double lo, double hi) {
    if (G->cfg->sampler == 0) {
        double sigma = pymax(sig_floor, (q[2] - q[0]) / 2.56);
        double y = q[1] + sigma * draw(rng, S_Z_YARDS, 1);
        return softclip(y, lo, hi);
    } else {
        double u = draw(rng, S_U_YQ, 0);
        double y = (u < 0.5) ? q[0] + (q[1] - q[0]) * (u / 0.5) : q[1] + (q[2] - q[1]) * ((u - 0.5) / 0.5);
        y = y + (0.0 + G->cfg->qy_noise * draw(rng, S_Z_YARDS, 1));
        /* np.clip(y, lo, hi) */
        double m = (y < lo) ? lo : y;
        return (m > hi) ? hi : m;
    }
}

/* P(pass) (FMC:407-427): heuristic, or softmax over play_model.xgb margins */
static double play_call_pass_prob(const FoGame *G, const FoState *s, int sd) {
    const FoConfig *c = G->cfg;
    if (c->policy == 0) return pass_prob_v1(s->down, s->dist, s->ytg, s->sec, sd);
    const FoForest *f = &g_forest[FO_PLAY_MODEL];
    int off = s->offense, de = off ^ 1;
    FoRow r;
    memset(&r, 0, sizeof(r));
    r.num[0] = (double)s->down; r.num[1] = s->dist; r.num[2] = s->ytg; r.num[3] = (s->ytg <= 20) ? 1.0 : 0.0;
    r.num[4] = (double)sd; r.num[5] = (double)s->sec; r.num[6] = 3.0; r.num[7] = 3.0;
    r.num[8] = c->sp[off][0]; r.num[9] = c->sp[off][1]; r.num[10] = c->sp[de][2]; r.num[11] = c->sp[de][0];
    /* play_model.json also reads goal_to_go, fourth_and_short, fg_range (features.pkl; FMC:1015-1017) */
    r.num[12] = (s->dist >= (s->ytg - 0.5)) ? 1.0 : 0.0;
    r.num[13] = (s->down == 4 && s->dist <= 2.0) ? 1.0 : 0.0;
    r.num[14] = (s->ytg <= 33) ? 1.0 : 0.0;
    r.num[15] = (s->sec > 1800) ? 1.0 : 2.0;
    r.num[16] = ((s->sec % 1800) <= 120) ? 1.0 : 0.0;
    for (int i = 0; i < f->n_scaled; ++i) {
        int k = f->scaler_cols[i];
        r.num[k] = (r.num[k] - f->scaler_mean[i]) / f->scaler_scale[i];
    }
    r.active[0] = c->coach_col[off]; r.active[1] = -1;
    double m[8]; float z[8], e[8];
    margins(f, &r, 0, -1, m);
    /* FMC:421-422: float32 softmax of margins / T */
    float T = (float)c->play_temp, zmax = 0.0f, sum = 0.0f;
    for (int k = 0; k < f->n_outputs; ++k) { z[k] = (float)m[k] / T; if (k == 0 || z[k] > zmax) zmax = z[k]; }
    for (int k = 0; k < f->n_outputs; ++k) { e[k] = expf_cr(z[k] - zmax); sum += e[k]; }
    double p = (double)(e[c->pass_class] / sum);
    return softclip(p, 0.02, 0.98);
}

/* FMC:1382-1421; returns 1 when a special-teams play consumed the iteration */
static int handle_fourth(FoGame *G, FoState *s, FoRng *rng) {
    if (s->down != 4) return 0;
    int team = s->offense;
    double ytg = s->ytg, dist = s->dist;
    int sd = s->score[team] - s->score[team ^ 1];
    double p_go = pymin(1.0, go_for_it_prob(ytg, dist, sd, s->sec) * 1.15);
    if (draw(rng, S_U_GO, 0) < p_go) { s->going = 1; G->n_go++; return 0; }
    if (ytg <= 38) {
        G->n_fga++;
        double p = field_goal_prob(ytg + 17);
        int good = draw(rng, S_U_FG, 0) < p;
        tick_clock(s, 12);
        if (good) { G->n_fg++; s->score[team] += 3; change_possession(s, 1, 75.0); }
        else change_possession(s, 1, 100.0 - ytg);
        return 1;
    }
    G->n_punt++;
    /* attempt_punt FMC:876-896 */
    double gross = pymax(30.0, 43.0 + 6.0 * draw(rng, S_Z_GROSS, 1));
    double ret = pymax(0.0, 6.0 + 3.0 * draw(rng, S_Z_RET, 1));
    double net = gross - ret;
    if (ytg <= 60) {
        double tb = softclip((60.0 - ytg) / 60.0, 0.10, 0.55);
        if (draw(rng, S_U_TB, 0) < tb) net = ytg - 25.0;
    }
    net = softclip(net, 15.0, ytg - 1.0);
    int inet = (int)net;                       /* int() truncates toward zero */
    tick_clock(s, 16);
    double spot = softclip(100.0 - (ytg - inet), 1, 99);
    change_possession(s, 1, spot);
    return 1;
}

/* FMC:1026-1257 */
static void simulate_play(FoGame *G, FoState *s, FoRng *rng) {
    if (s->sec <= 0) return;
    const FoConfig *c = G->cfg;
    int team = s->offense;
    int sd = s->score[team] - s->score[team ^ 1];
    double p_pass = play_call_pass_prob(G, s, sd);
    /* sample_categorical(["run","pass"], [1-p, p]) FMC:99-102 + Generator.choice */
    double a0 = 1.0 - p_pass, a1 = p_pass, sum = a0 + a1;
    a0 = a0 / sum; a1 = a1 / sum;
    double c0 = a0, c1 = a0 + a1;
    c0 = c0 / c1;
    int is_pass = !(draw(rng, S_U_CALL, 0) < c0);
    G->plays++;
    double ytg0 = s->ytg;
    double mz = G->mz[team];

    if (is_pass) {
        G->n_pass++;
        {
            const double u_qb = draw(rng, S_U_P1, 0);      /* sample_qb  FMC:625-627 */
            const double u_wr = draw(rng, S_U_WR, 0);      /* sample_target FMC:633-635 */
            if (G->usage) {
                G->cur[0] = sample_usage(G, team, 0, u_qb);
                G->cur[2] = sample_usage(G, team, 2, u_wr);
                credit(G, team, 2, 1, 1.0);                /* tgt += 1 (FMC:1074-1075) */
            }
        }
        FoRow r;
        fill_row(&r, G, s, sd, FO_PASS_STAGE1);
        double m1[1];
        margins(&g_forest[FO_PASS_STAGE1], &r, 0, -1, m1);
        double p1 = (double)sigmoid_f32((float)m1[0]);       /* inplace_predict -> probability */
        double p_complete = softclip(p1 + G->bias[team], 0.02, 0.98);
        if (draw(rng, S_U_COMP, 0) < p_complete) {
            G->n_comp++;
            double q[3];
            quants(G, s, sd, FO_PASS_YARDS, q);
            double yards = sample_yards(G, rng, q, 0.4, 0.0, s->ytg) * G->ymul[team];
            if (ytg0 > 25 && draw(rng, S_U_EX, 0) < 0.60 * explosive_prob(mz, ytg0)) {
                double ub = 0.35 + (0.95 - 0.35) * draw(rng, S_U_BOOST, 0);
                yards *= 1.0 + ub * (1.0 + 0.7 * mz);
                yards = pymin(yards, ytg0);
            }
            if (ytg0 <= 12 && s->down <= 3 &&
                draw(rng, S_U_FIN, 0) < rz_finish_prob_pass(ytg0, G->tanh35[team], s->down))
                yards = ytg0;
            if (G->usage) credit(G, team, 0, 1, 1.0);       /* att (FMC:1108-1109) */
            if (yards + 1e-9 >= s->ytg) {
                G->n_td++;
                if (G->usage) {                              /* FMC:1120-1127 */
                    credit(G, team, 0, 2, 1.0); credit(G, team, 0, 0, s->ytg); credit(G, team, 0, 3, 1.0);
                    credit(G, team, 2, 2, 1.0); credit(G, team, 2, 0, s->ytg); credit(G, team, 2, 3, 1.0);
                }
                s->score[team] += 7;
                s->going = 0;
                tick_clock(s, 20);
                change_possession(s, 1, 75.0);
            } else {
                if (G->usage) {                              /* FMC:1140-1145 */
                    credit(G, team, 0, 2, 1.0); credit(G, team, 0, 0, yards);
                    credit(G, team, 2, 2, 1.0); credit(G, team, 2, 0, yards);
                }
                s->going = 0;
                advance_down(s, yards);
                tick_clock(s, 26);
            }
            return;
        }
        /* stage 2 (FMC:751-770) */
        double raw[3];
        if (c->stage2_mode == 1) {
            FoRow r2; double m2[8]; float mf[3], pf[3];
            fill_row(&r2, G, s, sd, FO_PASS_STAGE2);
            margins(&g_forest[FO_PASS_STAGE2], &r2, 0, -1, m2);
            for (int k = 0; k < 3; ++k) mf[k] = (float)m2[k];
            softmax_f32(mf, 3, pf);
            for (int k = 0; k < 3; ++k) raw[k] = (double)pf[k];
        } else {
            for (int k = 0; k < 3; ++k) raw[k] = c->standin[k];
        }
        double p_inc = pymax(0.0, raw[0]), p_int = pymax(0.0, raw[1]), p_sck = pymax(0.0, raw[2]);
        p_sck *= 0.65;
        p_int = p_int * 1.20 + 0.004;
        double ssum = p_inc + p_int + p_sck;
        if (ssum == 0.0) ssum = 1.0;
        double b0 = p_inc / ssum, b1 = p_int / ssum, b2 = p_sck / ssum;
        double bs = (b0 + b1) + b2;
        b0 = b0 / bs; b1 = b1 / bs; b2 = b2 / bs;
        double d0 = b0, d1 = b0 + b1, d2 = (b0 + b1) + b2;
        d0 = d0 / d2; d1 = d1 / d2;
        double u2 = draw(rng, S_U_S2, 0);
        int outcome = (d0 <= u2) + (d1 <= u2);       /* searchsorted(cdf, u, 'right'), cdf[2] == 1 */
        if (outcome > 2) outcome = 2;
        if (outcome == 0) {                          /* incomplete FMC:1160-1168 */
            G->n_inc++;
            if (G->usage) credit(G, team, 0, 1, 1.0);       /* att (FMC:1163-1164) */
            s->down += 1;
            s->going = 0;
            tick_clock(s, 10);
        } else if (outcome == 2) {                   /* sack FMC:1170-1184 */
            G->n_sack++;
            if (G->usage) credit(G, team, 0, 5, 1.0);       /* sacks (FMC:1173-1174) */
            double q[3];
            quants(G, s, sd, FO_SACK_YARDS, q);
            double loss = -sample_yards(G, rng, q, 0.25, -20.0, 0.0);
            loss = pymax(0.0, loss);
            loss = pymin(loss, 100 - (100 - s->ytg));
            s->ytg += loss;
            s->dist += loss;
            s->down += 1;
            s->going = 0;
            tick_clock(s, 24);
        } else {                                     /* intercepted FMC:1186-1199 */
            G->n_int++;
            if (G->usage) { credit(G, team, 0, 1, 1.0); credit(G, team, 0, 4, 1.0); }   /* att, INT (FMC:1190-1192) */
            double ret = softclip(6 + 5 * draw(rng, S_Z_INT, 1), 0, s->ytg);
            double spot = 100.0 - (s->ytg - ret);
            s->going = 0;
            change_possession(s, 1, spot);
            tick_clock(s, 12);
        }
        return;
    }
    /* run FMC:1201-1257 */
    G->n_run++;
    {
        const double u_rb = draw(rng, S_U_P1, 0);    /* sample_rusher FMC:629-631 */
        if (G->usage) {
            G->cur[1] = sample_usage(G, team, 1, u_rb);
            credit(G, team, 1, 1, 1.0);              /* att (FMC:1206-1208) */
        }
    }
    double q[3];
    quants(G, s, sd, FO_RUN_YARDS, q);
    double yards = sample_yards(G, rng, q, 0.35, -4.0, s->ytg) * G->ymul[team];
    if (ytg0 > 25 && draw(rng, S_U_EX, 0) < 0.5 * explosive_prob(mz, ytg0)) {
        double ub = 0.2 + (0.5 - 0.2) * draw(rng, S_U_BOOST, 0);
        yards *= 1.0 + ub * (1.0 + 0.6 * mz);
        yards = pymin(yards, ytg0);
    }
    if (ytg0 <= 9 && s->down <= 3) {
        if (draw(rng, S_U_FIN, 0) < rz_finish_prob_run(ytg0, G->tanh35[team], s->down)) yards = ytg0;
    }
    if (yards + 1e-9 >= ytg0) {
        G->n_td++;
        if (G->usage) { credit(G, team, 1, 0, s->ytg); credit(G, team, 1, 3, 1.0); }   /* FMC:1232-1234 */
        s->score[team] += 7;
        tick_clock(s, 28);
        change_possession(s, 1, 75.0);
        s->going = 0;
    } else {
        if (G->usage) credit(G, team, 1, 0, yards);   /* FMC:1246-1247 */
        advance_down(s, yards);
        tick_clock(s, 28);
        s->going = 0;
    }
}

static void game_constants(FoGame *G, const FoConfig *c) {
    memset(G, 0, sizeof(*G));
    G->cfg = c;
    for (int off = 0; off < 2; ++off) {
        double O = c->sp[off][1], D = c->sp[off ^ 1][2];
        G->bias[off] = 0.12 * (O - D) / 40.0;                 /* matchup_bias FMC:431-433 */
        G->ymul[off] = 1.0 + 0.10 * tanh((O - D) / 30.0);     /* yardage_multiplier FMC:435-437 */
        G->mz[off] = (O - D) / 40.0;                          /* mismatch_z FMC:440-442 */
        G->tanh35[off] = tanh((O - D) / 35.0);                /* FMC:448, 456 */
    }
}

/* FMC:1428-1464.  first = team index that receives the opening kickoff. */
static void simulate_game(FoGame *G, FoRng *rng, int first, int score_out[2], int *iters_out, double *trace) {
    FoState s;
    s.offense = first; s.sec = 3600; s.down = 1; s.dist = 10.0; s.ytg = 75.0; s.period = 1; s.going = 0;
    s.score[0] = s.score[1] = 0;
    int it = 0;
    while (s.sec > 0) {
        rng->iter = it;
        rng->have[0] = rng->have[1] = rng->have[2] = rng->have[3] = 0;
        if (trace && it < FO_MAX_ITERS) {
            double *t = trace + (size_t)it * FO_TRACE_COLS;
            t[0] = (s.offense == first) ? 1.0 : 0.0; t[1] = s.down; t[2] = s.sec;
            t[3] = s.score[first]; t[4] = s.score[first ^ 1]; t[5] = s.dist; t[6] = s.ytg; t[7] = s.going;
        }
        ++it;
        if (handle_fourth(G, &s, rng)) continue;
        simulate_play(G, &s, rng);
    }
    score_out[0] = s.score[0]; score_out[1] = s.score[1];
    *iters_out = it;
    G->iters += it;
}

/*
 * Simulate games [game0, game0+n) of one matchup.  Game g: team (g & 1) receives the opening
 * kickoff (pairs A-first / B-first, FMC:1321-1328, 1502-1503).
 *   rng_mode 0: stream[n][360][16] injected draws;  1: Philox keyed (seed, matchup, game id).
 *   scores[n][2] = points of team 0 / team 1;  iters[n];  trace[n][360][8] or NULL.
 *   counters[16]: plays, iters, pass, comp, inc, int, sack, run, td, fga, fg, punt, go (summed).
 */
static int simulate_range(const FoConfig *cfg, const FoTeamUsage *usage, int n_slots, double *players,
                          long n, long game0, int matchup, int rng_mode, const double *stream,
                          uint64_t seed, int *scores, int *iters, double *trace, long *counters, int n_threads) {
    for (int m = FO_PASS_STAGE1; m <= FO_SACK_YARDS; ++m) {
        if (m == FO_PASS_STAGE2 && cfg->stage2_mode == 0) continue;
        if (!g_forest[m].loaded) return -2;
    }
    if (cfg->policy == 1 && !g_forest[FO_PLAY_MODEL].loaded) return -2;
    long tot[16];
    memset(tot, 0, sizeof(tot));
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        FoGame G;
        game_constants(&G, cfg);
        if (usage) {
            G.usage = usage;
            G.n_slots = n_slots;
            for (int t = 0; t < 2; ++t)
                for (int r = 0; r < 3; ++r) {
                    const FoUsage *u = &usage[t].role[r];
                    double acc = 0.0;
                    for (int i = 0; i < u->n; ++i) { acc += u->share[i]; G.cdf[t][r][i] = acc; }   /* np.cumsum */
                    for (int i = 0; i < u->n; ++i) G.cdf[t][r][i] /= acc;                           /* cdf /= cdf[-1] */
                }
        }
#pragma omp for schedule(dynamic, 16)
        for (long i = 0; i < n; ++i) {
            long g = game0 + i;
            FoRng rng;
            memset(&rng, 0, sizeof(rng));
            rng.mode = rng_mode;
            rng.rec = stream ? stream + (size_t)i * FO_MAX_ITERS * FO_N_SLOTS : NULL;
            rng.key[0] = (uint32_t)seed; rng.key[1] = (uint32_t)(seed >> 32);
            rng.ctr_game[0] = (uint32_t)((uint64_t)g); rng.ctr_game[1] = (uint32_t)((uint64_t)g >> 32);
            rng.ctr_game[2] = (uint32_t)matchup;
            int it = 0;
            G.box = (usage && players) ? players + (size_t)i * 2 * n_slots * FO_PLAYER_FIELDS : NULL;
            G.cur[0] = G.cur[1] = G.cur[2] = 0;
            simulate_game(&G, &rng, (int)(g & 1), scores + 2 * i, &it,
                          trace ? trace + (size_t)i * FO_MAX_ITERS * FO_TRACE_COLS : NULL);
            if (iters) iters[i] = it;
        }
#pragma omp critical
        {
            tot[0] += G.plays; tot[1] += G.iters; tot[2] += G.n_pass; tot[3] += G.n_comp; tot[4] += G.n_inc;
            tot[5] += G.n_int; tot[6] += G.n_sack; tot[7] += G.n_run; tot[8] += G.n_td; tot[9] += G.n_fga;
            tot[10] += G.n_fg; tot[11] += G.n_punt; tot[12] += G.n_go;
        }
    }
    if (counters) memcpy(counters, tot, sizeof(tot));
    return 0;
}

int fo_simulate(const FoConfig *cfg, long n, long game0, int matchup, int rng_mode, const double *stream,
                uint64_t seed, int *scores, int *iters, double *trace, long *counters, int n_threads) {
    return simulate_range(cfg, NULL, 0, NULL, n, game0, matchup, rng_mode, stream, seed, scores, iters, trace,
                          counters, n_threads);
}

/*
 * Same with usage tables (FMC:625-635) and the per-player box of the focus names (FMC:1259-1299):
 *   usage[2]   team A / team B;  players[n][2][n_slots][FO_PLAYER_FIELDS] (zeroed by the caller), fields
 *   yds, att|tgt, comp|rec, td, INT, sacks of the name tracked in that slot.
 */
int fo_simulate_players(const FoConfig *cfg, const FoTeamUsage *usage, int n_slots, double *players,
                        long n, long game0, int matchup, int rng_mode, const double *stream,
                        uint64_t seed, int *scores, int *iters, double *trace, long *counters, int n_threads) {
    if (!usage || n_slots < 0) return -3;
    for (int t = 0; t < 2; ++t)
        for (int r = 0; r < 3; ++r) {
            const FoUsage *u = &usage[t].role[r];
            if (u->n < 1 || u->n > FO_MAX_USAGE) return -3;
            for (int i = 0; i < u->n; ++i)
                if (u->slot[i] >= n_slots) return -3;
        }
    return simulate_range(cfg, usage, n_slots, players, n, game0, matchup, rng_mode, stream, seed, scores, iters,
                          trace, counters, n_threads);
}

int fo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- scalar hooks so tests can pin the closed-form helpers against tests/golden/ref_scalars.json ---- */
double fo_pass_prob_v1(int down, double distance, double ytg, int sec, int sd) { return pass_prob_v1(down, distance, ytg, sec, sd); }
double fo_go_for_it_prob(double ytg, double dist, int sd, int sec) { return go_for_it_prob(ytg, dist, sd, sec); }
double fo_field_goal_prob(double d) { return field_goal_prob(d); }
/* out = {matchup_bias, yardage_multiplier, mismatch_z, explosive_prob(ytg), rz_finish_prob_pass, rz_finish_prob_run} */
/* P(pass) of play_call_pass_prob_binary (FMC:407-427) for one state; `off` = team index on offense */
double fo_play_pass_prob(const FoConfig *cfg, int off, int down, double dist, double ytg, int sd, int sec) {
    FoGame G;
    game_constants(&G, cfg);
    FoState s;
    memset(&s, 0, sizeof(s));
    s.offense = off; s.down = down; s.dist = dist; s.ytg = ytg; s.sec = sec;
    return play_call_pass_prob(&G, &s, sd);
}

void fo_modifiers(double off_offense, double def_defense, double ytg, int down, double *out) {
    FoConfig c;
    FoGame G;
    memset(&c, 0, sizeof(c));
    c.sp[0][1] = off_offense; c.sp[1][2] = def_defense;
    game_constants(&G, &c);
    out[0] = G.bias[0]; out[1] = G.ymul[0]; out[2] = G.mz[0];
    out[3] = explosive_prob(G.mz[0], ytg);
    out[4] = rz_finish_prob_pass(ytg, G.tanh35[0], down);
    out[5] = rz_finish_prob_run(ytg, G.tanh35[0], down);
}
