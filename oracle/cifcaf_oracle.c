/*
 * oracle/cifcaf_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference CifCaf decoder (openpifpaf csrc),
 * written from the reference's behaviour, one function per reference function.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path
 * (openpifpaf_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  This restatement is checked bit-for-bit against the
 * unmodified reference C++ compiled from /root/reference into oracle/_ref/
 * (oracle/build_ref.py) by tests/test_oracle.py, and against the golden
 * vectors under tests/golden/ that oracle/make_golden.py dumped from that
 * reference build.  (The reference's own tests hold no numeric KAT for the
 * decoder)
 *
 * All citations are relative to /root/reference/src/openpifpaf/csrc/.
 * Numerics: every expression keeps the reference's float/double promotions
 * (SURVEY.md appendix B).  Compile WITHOUT -ffast-math / -march=native so no
 * FMA contraction happens (the reference build has none either).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    /* CifHr statics: src/cif_hr.cpp:13-15 */
    int64_t cifhr_neighbors;          /* 16 */
    double cifhr_threshold;           /* 0.3 */
    int32_t cifhr_ablation_skip;      /* 0 */
    /* CifSeeds statics: src/cif_seeds.cpp:11-14 */
    double seed_threshold;            /* 0.2 */
    int32_t seeds_ablation_nms;       /* 0 */
    int32_t seeds_ablation_no_rescore;/* 0 */
    /* CafScored: src/caf_scored.cpp:11-12, cifcaf.cpp:153 */
    double caf_score_th;              /* default_score_th 0.3 */
    double caf_cif_floor;             /* 0.1 */
    int32_t caf_ablation_no_rescore;  /* 0 */
    /* CifCaf statics: src/cifcaf.cpp:18-24 */
    int32_t block_joints;             /* 0 (no effect, cifcaf.cpp:291-295) */
    int32_t greedy;                   /* 0 */
    double keypoint_threshold;        /* 0.15 */
    double keypoint_threshold_rel;    /* 0.5 */
    int32_t reverse_match;            /* 1 */
    int32_t force_complete;           /* 0 */
    double force_complete_caf_th;     /* 0.001 */
    /* NMSKeypoints statics: src/nms_keypoints.cpp:12-14 */
    double nms_suppression;           /* 1e-5 */
    double nms_instance_threshold;    /* 0.15 */
    double nms_keypoint_threshold;    /* 0.15 */
    /* Occupancy(2.0, 4.0): include/openpifpaf/decoder/cifcaf.hpp:103 */
    double occ_reduction;             /* 2.0 */
    double occ_min_scale;             /* 4.0 */
    /* CifHr revision of the call; a fresh reference instance has 1.0
     * (src/cif_hr.cpp:115: revision++ from 0.0). */
    double cifhr_revision;            /* 1.0 */
    /* Tie order of the seed sort (src/cif_seeds.cpp:94-97 uses std::sort, which is
     * unstable): 0 = libstdc++ introsort restated (matches oracle/_ref bit for bit),
     * 1 = stable by fill order (f, j, i) -- the documented contract of the CUDA path. */
    int32_t seed_sort_stable;         /* 0 */
} oracle_params_t;

void oracle_default_params(oracle_params_t* p) {
    memset(p, 0, sizeof(*p));
    p->cifhr_neighbors = 16;
    p->cifhr_threshold = 0.3;
    p->seed_threshold = 0.2;
    p->caf_score_th = 0.3;
    p->caf_cif_floor = 0.1;
    p->keypoint_threshold = 0.15;
    p->keypoint_threshold_rel = 0.5;
    p->reverse_match = 1;
    p->force_complete_caf_th = 0.001;
    p->nms_suppression = 0.00001;
    p->nms_instance_threshold = 0.15;
    p->nms_keypoint_threshold = 0.15;
    p->occ_reduction = 2.0;
    p->occ_min_scale = 4.0;
    p->cifhr_revision = 1.0;
}

static inline int64_t clamp_i64(int64_t v, int64_t lo, int64_t hi) {
    /* std::clamp(v, lo, hi) */
    return v < lo ? lo : (hi < v ? hi : v);
}

/* ------------------------------------------------------------------ CifHr */

/* src/cif_hr.cpp:18-25 */
static inline float approx_exp(float x) {
    if (x > 2.0 || x < -2.0) return 0.0f;
    x = (float)(1.0 + (double)x / 8.0);
    x *= x;
    x *= x;
    x *= x;
    return x;
}

/* src/cif_hr.cpp:58-89. acc is [F][H][W] contiguous. */
static void cifhr_add_gauss(float* acc, int64_t H, int64_t W, double revision,
                            int64_t f, float v, float x, float y, float sigma, float truncate) {
    int64_t minx = clamp_i64((int64_t)(x - truncate * sigma), 0, W - 1);
    int64_t miny = clamp_i64((int64_t)(y - truncate * sigma), 0, H - 1);
    int64_t maxx = clamp_i64((int64_t)(x + truncate * sigma + 1), minx + 1, W);
    int64_t maxy = clamp_i64((int64_t)(y + truncate * sigma + 1), miny + 1, H);

    float sigma2 = sigma * sigma;
    float truncate2_sigma2 = truncate * truncate * sigma2;
    float revision_f = (float)revision;
    float revision_p1_f = (float)(revision + 1.0);
    for (int64_t xx = minx; xx < maxx; xx++) {
        float deltax2 = ((float)xx - x) * ((float)xx - x);
        for (int64_t yy = miny; yy < maxy; yy++) {
            float deltay2 = ((float)yy - y) * ((float)yy - y);
            if (deltax2 + deltay2 > truncate2_sigma2) continue;
            float vv;
            if (deltax2 < 0.25 && deltay2 < 0.25) {
                vv = v;
            } else {
                vv = v * approx_exp((float)(-0.5 * (double)(deltax2 + deltay2) / (double)sigma2));
            }
            float* entry = &acc[(f * H + yy) * W + xx];
            *entry = fmaxf(*entry, revision_f) + vv;
            *entry = fminf(*entry, revision_p1_f);
        }
    }
}

/* src/cif_hr.cpp:28-55.  cif is [F][5][h][w]; acc is [F][H][W] with
 * H=(h-1)*stride+1, W=(w-1)*stride+1 (src/cif_hr.cpp:110-114), pre-zeroed for a
 * fresh instance. */
void oracle_cifhr_accumulate(const float* cif, int64_t F, int64_t h, int64_t w, int64_t stride,
                             double min_scale, double factor, const oracle_params_t* p,
                             float* acc) {
    if (p->cifhr_ablation_skip) return;
    int64_t H = (h - 1) * stride + 1, W = (w - 1) * stride + 1;
    float min_scale_f = (float)(min_scale / (double)stride);
    int64_t hw = h * w;
    for (int64_t f = 0; f < F; f++) {
        const float* cf = cif + f * 5 * hw;
        for (int64_t j = 0; j < h; j++) {
            for (int64_t i = 0; i < w; i++) {
                float v = cf[1 * hw + j * w + i];
                if ((double)v < p->cifhr_threshold) continue;
                float scale = cf[4 * hw + j * w + i];
                if (scale < min_scale_f) continue;
                float x = cf[2 * hw + j * w + i] * (float)stride;
                float y = cf[3 * hw + j * w + i] * (float)stride;
                float sigma = fmaxf(1.0f, (float)(0.5 * (double)scale * (double)stride));
                float vn = (float)((double)(v / (float)p->cifhr_neighbors) * factor);
                cifhr_add_gauss(acc, H, W, p->cifhr_revision, f, vn, x, y, sigma, 1.0f);
            }
        }
    }
}

/* src/cif_seeds.cpp:17-30 and src/caf_scored.cpp:15-26 (identical bodies). */
static float cifhr_value(const float* acc, int64_t F, int64_t H, int64_t W, double revision,
                         int64_t f, float x, float y, float default_value) {
    float max_x = (float)((double)(float)W - 0.51);
    float max_y = (float)((double)(float)H - 0.51);
    if (f >= F || x < -0.49 || y < -0.49 || x > max_x || y > max_y) return default_value;
    float value = (float)((double)acc[(f * H + (int64_t)((double)y + 0.5)) * W
                                      + (int64_t)((double)x + 0.5)] - revision);
    if (value < 0.0) return default_value;
    return value;
}

/* ------------------------------------------------------------------ CifSeeds */

typedef struct { int64_t f; float v, x, y, s; int64_t order; float h; } seed_t;   /* s doubles as DetSeed::w, h = DetSeed::h */

/* libstdc++ std::sort (bits/stl_algo.h: __introsort_loop + __final_insertion_sort,
 * _S_threshold = 16) restated for comp(a,b) = a.v > b.v, so that exact float ties
 * land in the same order as in the compiled reference. */
#define SEED_COMP(a, b) ((a).v > (b).v)
static void seed_swap(seed_t* a, seed_t* b) { seed_t t = *a; *a = *b; *b = t; }
static void seed_heap_adjust(seed_t* first, int64_t hole, int64_t len, seed_t value) {
    int64_t top = hole, second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (SEED_COMP(first[second], first[second - 1])) second--;
        first[hole] = first[second]; hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        first[hole] = first[second - 1]; hole = second - 1;
    }
    int64_t parent = (hole - 1) / 2;
    while (hole > top && SEED_COMP(first[parent], value)) {
        first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void seed_heapsort(seed_t* first, int64_t len) {
    /* std::__partial_sort(first, last, last): __heap_select (make_heap) + __sort_heap */
    if (len >= 2) {
        int64_t parent = (len - 2) / 2;
        for (;;) {
            seed_t value = first[parent];
            seed_heap_adjust(first, parent, len, value);
            if (parent == 0) break;
            parent--;
        }
    }
    int64_t last = len;
    while (last > 1) {
        --last;
        seed_t value = first[last];
        first[last] = first[0];
        seed_heap_adjust(first, 0, last, value);
    }
}
static void seed_introsort_loop(seed_t* first, seed_t* last, int64_t depth_limit) {
    while (last - first > 16) {
        if (depth_limit == 0) { seed_heapsort(first, last - first); return; }
        --depth_limit;
        seed_t* mid = first + (last - first) / 2;
        seed_t *a = first + 1, *b = mid, *c = last - 1;
        if (SEED_COMP(*a, *b)) {
            if (SEED_COMP(*b, *c)) seed_swap(first, b);
            else if (SEED_COMP(*a, *c)) seed_swap(first, c);
            else seed_swap(first, a);
        } else if (SEED_COMP(*a, *c)) seed_swap(first, a);
        else if (SEED_COMP(*b, *c)) seed_swap(first, c);
        else seed_swap(first, b);
        seed_t *lo = first + 1, *hi = last;
        for (;;) {
            while (SEED_COMP(*lo, *first)) ++lo;
            --hi;
            while (SEED_COMP(*first, *hi)) --hi;
            if (!(lo < hi)) break;
            seed_swap(lo, hi);
            ++lo;
        }
        seed_introsort_loop(lo, last, depth_limit);
        last = lo;
    }
}
static void seed_unguarded_linear_insert(seed_t* last) {
    seed_t val = *last; seed_t* next = last - 1;
    while (SEED_COMP(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void seed_insertion_sort(seed_t* first, seed_t* last) {
    if (first == last) return;
    for (seed_t* i = first + 1; i != last; ++i) {
        if (SEED_COMP(*i, *first)) {
            seed_t val = *i;
            memmove(first + 1, first, sizeof(seed_t) * (size_t)(i - first));
            *first = val;
        } else seed_unguarded_linear_insert(i);
    }
}
static void seed_std_sort(seed_t* first, int64_t n) {
    if (n <= 1) return;
    int64_t lg = 0; for (int64_t m = n; m > 1; m >>= 1) lg++;
    seed_introsort_loop(first, first + n, 2 * lg);
    if (n > 16) {
        seed_insertion_sort(first, first + 16);
        for (seed_t* i = first + 16; i != first + n; ++i) seed_unguarded_linear_insert(i);
    } else seed_insertion_sort(first, first + n);
}

static int seed_cmp(const void* a, const void* b) {
    /* stable mode: v descending, exact ties by fill order (f, j, i). */
    const seed_t* sa = (const seed_t*)a; const seed_t* sb = (const seed_t*)b;
    if (sa->v > sb->v) return -1;
    if (sa->v < sb->v) return 1;
    return sa->order < sb->order ? -1 : (sa->order > sb->order ? 1 : 0);
}

/* src/cif_seeds.cpp:33-66 + 93-114.  Returns n_seeds; writes up to cap. */
int64_t oracle_cifseeds(const float* cif, int64_t F, int64_t h, int64_t w, int64_t stride,
                        const float* acc, const oracle_params_t* p,
                        int64_t* out_f, float* out_vxys, int64_t cap) {
    int64_t H = (h - 1) * stride + 1, W = (w - 1) * stride + 1;
    int64_t hw = h * w;
    seed_t* seeds = (seed_t*)malloc(sizeof(seed_t) * (size_t)(F * hw + 1));
    int64_t n = 0;
    for (int64_t f = 0; f < F; f++) {
        const float* cf = cif + f * 5 * hw;
        for (int64_t j = 0; j < h; j++) {
            for (int64_t i = 0; i < w; i++) {
                float c = cf[1 * hw + j * w + i];
                if ((double)c < p->seed_threshold) continue;
                if (p->seeds_ablation_nms) {
                    /* torch.max_pool2d(confidence, 3, 1, 1): src/cif_seeds.cpp:36-40,49-51 */
                    float m = c;
                    for (int64_t dj = -1; dj <= 1; dj++) for (int64_t di = -1; di <= 1; di++) {
                        int64_t jj = j + dj, ii = i + di;
                        if (jj < 0 || jj >= h || ii < 0 || ii >= w) continue;
                        float o = cf[1 * hw + jj * w + ii];
                        if (o > m) m = o;
                    }
                    if (c < m) continue;
                }
                float x = cf[2 * hw + j * w + i] * (float)stride;
                float y = cf[3 * hw + j * w + i] * (float)stride;
                if (!p->seeds_ablation_no_rescore) {
                    c = (float)(0.9 * (double)cifhr_value(acc, F, H, W, p->cifhr_revision, f, x, y, -1.0f)
                                + 0.1 * (double)c);
                }
                if ((double)c < p->seed_threshold) continue;
                float s = cf[4 * hw + j * w + i] * (float)stride;
                seeds[n].f = f; seeds[n].v = c; seeds[n].x = x; seeds[n].y = y; seeds[n].s = s;
                seeds[n].order = n;
                n++;
            }
        }
    }
    if (p->seed_sort_stable) qsort(seeds, (size_t)n, sizeof(seed_t), seed_cmp);
    else seed_std_sort(seeds, n);
    for (int64_t k = 0; k < n && k < cap; k++) {
        out_f[k] = seeds[k].f;
        out_vxys[4 * k + 0] = seeds[k].v; out_vxys[4 * k + 1] = seeds[k].x;
        out_vxys[4 * k + 2] = seeds[k].y; out_vxys[4 * k + 3] = seeds[k].s;
    }
    free(seeds);
    return n;
}

/* ------------------------------------------------------------------ CafScored */

/* src/caf_scored.cpp:29-83.  caf is [C][8][h][w]; skeleton is [C][2] 0-based.
 * fwd/bwd are [C][h*w][7] capacity buffers; n_fwd/n_bwd are [C]. */
void oracle_cafscored(const float* caf, int64_t C, int64_t h, int64_t w, int64_t stride,
                      const int64_t* skeleton, const float* acc, int64_t F, int64_t H, int64_t W,
                      double score_th_arg, const oracle_params_t* p,
                      float* fwd, int64_t* n_fwd, float* bwd, int64_t* n_bwd) {
    /* include/openpifpaf/decoder/utils/caf_scored.hpp:56-57 */
    double score_th = score_th_arg >= 0.0 ? score_th_arg : p->caf_score_th;
    double cif_floor = p->caf_cif_floor;
    int64_t hw = h * w;
    for (int64_t f = 0; f < C; f++) {
        const float* cf = caf + f * 8 * hw;
        float* ff = fwd + f * hw * 7; float* bb = bwd + f * hw * 7;
        int64_t nf = 0, nb = 0;
        for (int64_t j = 0; j < h; j++) {
            for (int64_t i = 0; i < w; i++) {
                int64_t o = j * w + i;
                float c = cf[1 * hw + o];
                if ((double)c < score_th) continue;
                float x1 = cf[2 * hw + o] * (float)stride;
                float y1 = cf[3 * hw + o] * (float)stride;
                float x2 = cf[4 * hw + o] * (float)stride;
                float y2 = cf[5 * hw + o] * (float)stride;
                float s1 = cf[6 * hw + o] * (float)stride;
                float s2 = cf[7 * hw + o] * (float)stride;
                float cfw = c, cbw = c;
                if (!p->caf_ablation_no_rescore) {
                    float forward_hr = cifhr_value(acc, F, H, W, p->cifhr_revision, skeleton[2 * f + 1], x2, y2, 0.0f);
                    float backward_hr = cifhr_value(acc, F, H, W, p->cifhr_revision, skeleton[2 * f + 0], x1, y1, 0.0f);
                    cfw = (float)((double)c * (cif_floor + (1.0 - cif_floor) * (double)forward_hr));
                    cbw = (float)((double)c * (cif_floor + (1.0 - cif_floor) * (double)backward_hr));
                }
                if ((double)cfw > score_th) {
                    float* e = ff + 7 * nf++;
                    e[0] = cfw; e[1] = x1; e[2] = y1; e[3] = x2; e[4] = y2; e[5] = s1; e[6] = s2;
                }
                if ((double)cbw > score_th) {
                    float* e = bb + 7 * nb++;
                    e[0] = cbw; e[1] = x2; e[2] = y2; e[3] = x1; e[4] = y1; e[5] = s2; e[6] = s1;
                }
            }
        }
        n_fwd[f] = nf; n_bwd[f] = nb;
    }
}

/* ------------------------------------------------------------------ Occupancy */

typedef struct {
    uint8_t* occ; int64_t F, H, W;   /* [F][H][W] bytes, 1 = occupied */
    double reduction, min_scale_reduced;
} occupancy_t;

/* src/occupancy.cpp:46-68 (shape) */
static void occupancy_reset(occupancy_t* o, int64_t F, int64_t Hhr, int64_t Whr, const oracle_params_t* p) {
    o->reduction = p->occ_reduction;
    o->min_scale_reduced = p->occ_min_scale / p->occ_reduction;
    o->F = F;
    o->H = (int64_t)((double)Hhr / o->reduction) + 1;
    o->W = (int64_t)((double)Whr / o->reduction) + 1;
    o->occ = (uint8_t*)calloc((size_t)(F * o->H * o->W), 1);
}
/* src/occupancy.cpp:71-77: clear == revision++ == everything unoccupied */
static void occupancy_clear(occupancy_t* o) { memset(o->occ, 0, (size_t)(o->F * o->H * o->W)); }

/* src/occupancy.cpp:13-29 */
static void occupancy_set(occupancy_t* o, int64_t f, double x, double y, double sigma) {
    if (o->reduction != 1.0) {
        x /= o->reduction; y /= o->reduction;
        sigma = fmax(o->min_scale_reduced, sigma / o->reduction);
    }
    int64_t minx = clamp_i64((int64_t)(x - sigma), 0, o->W - 1);
    int64_t miny = clamp_i64((int64_t)(y - sigma), 0, o->H - 1);
    int64_t maxx = clamp_i64((int64_t)(x + sigma), minx + 1, o->W);
    int64_t maxy = clamp_i64((int64_t)(y + sigma), miny + 1, o->H);
    for (int64_t yy = miny; yy < maxy; yy++)
        memset(o->occ + (f * o->H + yy) * o->W + minx, 1, (size_t)(maxx - minx));
}
/* src/occupancy.cpp:32-43 */
static int occupancy_get(const occupancy_t* o, int64_t f, double x, double y) {
    if (f >= o->F) return 1;
    if (o->reduction != 1.0) { x /= o->reduction; y /= o->reduction; }
    int64_t xi = clamp_i64((int64_t)x, 0, o->W - 1);
    int64_t yi = clamp_i64((int64_t)y, 0, o->H - 1);
    return o->occ[(f * o->H + yi) * o->W + xi] != 0;
}

/* ------------------------------------------------------------------ grow */

typedef struct { double v, x, y, s; } joint_t;   /* include/.../cifcaf.hpp:21-28 */

/* src/cifcaf.cpp:32-103.  caf is [n][7]. */
static joint_t grow_connection_blend(const float* caf, int64_t n, double x, double y,
                                     double xy_scale, double filter_sigmas, int only_max) {
    joint_t zero = {0, 0, 0, 0};
    xy_scale = fmax(xy_scale, 0.5);
    float sigma_filter = (float)(filter_sigmas * xy_scale / 2.0);
    float sigma2 = (float)(0.25 * xy_scale * xy_scale);
    int64_t score_1_i = 0, score_2_i = 0;
    float score_1 = 0.0f, score_2 = 0.0f;
    for (int64_t i = 0; i < n; i++) {
        const float* e = caf + 7 * i;
        if ((double)e[1] < x - (double)sigma_filter) continue;
        if ((double)e[1] > x + (double)sigma_filter) continue;
        if ((double)e[2] < y - (double)sigma_filter) continue;
        if ((double)e[2] > y + (double)sigma_filter) continue;
        double dx = (double)e[1] - x, dy = (double)e[2] - y;
        float d2 = (float)(dx * dx + dy * dy);          /* std::pow(.,2) is exact x*x */
        float score = (float)(exp(-0.5 * (double)d2 / (double)sigma2) * (double)e[0]);
        if (score >= score_1) {
            score_2_i = score_1_i; score_2 = score_1;
            score_1_i = i; score_1 = score;
        } else if (score > score_2) {
            score_2_i = i; score_2 = score;
        }
    }
    if (score_1 == 0.0) return zero;

    const float* e1 = caf + 7 * score_1_i;
    float entry_1[3] = { e1[3], e1[4], fmaxf(0.0f, e1[6]) };
    if (only_max) {
        joint_t r = { score_1, entry_1[0], entry_1[1], entry_1[2] };
        return r;
    }
    if (score_2 < 0.01 || (double)score_2 < 0.5 * (double)score_1) {
        joint_t r = { 0.5 * (double)score_1, entry_1[0], entry_1[1], entry_1[2] };
        return r;
    }
    const float* e2 = caf + 7 * score_2_i;
    float entry_2[3] = { e2[3], e2[4], fmaxf(0.0f, e2[6]) };
    double bdx = (double)(entry_1[0] - entry_2[0]), bdy = (double)(entry_1[1] - entry_2[1]);
    float blend_d2 = (float)(bdx * bdx + bdy * bdy);
    if ((double)blend_d2 > ((double)entry_1[2] * (double)entry_1[2]) / 4.0) {
        joint_t r = { 0.5 * (double)score_1, entry_1[0], entry_1[1], entry_1[2] };
        return r;
    }
    joint_t r = {
        0.5 * (double)(score_1 + score_2),
        (score_1 * entry_1[0] + score_2 * entry_2[0]) / (score_1 + score_2),
        (score_1 * entry_1[1] + score_2 * entry_2[1]) / (score_1 + score_2),
        (score_1 * entry_1[2] + score_2 * entry_2[2]) / (score_1 + score_2)
    };
    return r;
}

/* exported for the free-op parity test (module.cpp:60, cifcaf.cpp:105-113): returns x,y,s,v */
void oracle_grow_connection_blend(const float* caf, int64_t n, double x, double y, double s,
                                  double filter_sigmas, int only_max, double* out_xysv) {
    joint_t j = grow_connection_blend(caf, n, x, y, s, filter_sigmas, only_max);
    out_xysv[0] = j.x; out_xysv[1] = j.y; out_xysv[2] = j.s; out_xysv[3] = j.v;
}

typedef struct { float max_score; joint_t joint; int64_t start_i, end_i; } frontier_entry_t;

typedef struct {
    int64_t K, C;
    const int64_t* skeleton;
    const oracle_params_t* p;
    int64_t occ_n_fields;
    /* std::priority_queue<FrontierEntry, vector, FrontierCompare> restated with
     * libstdc++'s push_heap/pop_heap (bits/stl_heap.h) so tie order matches. */
    frontier_entry_t* heap; int64_t heap_n;
    uint8_t* in_frontier;   /* [K][K] */
} grow_ctx_t;

static inline int frontier_less(const frontier_entry_t* a, const frontier_entry_t* b) {
    return a->max_score < b->max_score;   /* src/cifcaf.cpp:27-29 */
}
static void heap_push_up(frontier_entry_t* first, int64_t hole, int64_t top, frontier_entry_t value) {
    int64_t parent = (hole - 1) / 2;
    while (hole > top && frontier_less(&first[parent], &value)) {
        first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void frontier_push(grow_ctx_t* g, frontier_entry_t e) {
    g->heap[g->heap_n++] = e;
    heap_push_up(g->heap, g->heap_n - 1, 0, e);
}
static frontier_entry_t frontier_pop(grow_ctx_t* g) {
    frontier_entry_t top = g->heap[0];
    if (g->heap_n > 1) {
        int64_t len = g->heap_n - 1;
        frontier_entry_t value = g->heap[len];
        g->heap[len] = g->heap[0];
        int64_t hole = 0, second = 0;
        while (second < (len - 1) / 2) {
            second = 2 * (second + 1);
            if (frontier_less(&g->heap[second], &g->heap[second - 1])) second--;
            g->heap[hole] = g->heap[second]; hole = second;
        }
        if ((len & 1) == 0 && second == (len - 2) / 2) {
            second = 2 * (second + 1);
            g->heap[hole] = g->heap[second - 1]; hole = second - 1;
        }
        heap_push_up(g->heap, hole, 0, value);
    }
    g->heap_n--;
    return top;
}

/* src/cifcaf.cpp:316-346 */
static void frontier_add_from(grow_ctx_t* g, const joint_t* joints, int64_t start_i) {
    float max_score = (float)sqrt(joints[start_i].v);
    for (int64_t f = 0; f < g->C; f++) {
        int64_t pair_0 = g->skeleton[2 * f], pair_1 = g->skeleton[2 * f + 1];
        if (pair_0 == start_i) {
            if (joints[pair_1].v > 0.0) continue;
            if (g->in_frontier[pair_0 * g->K + pair_1]) continue;
            frontier_entry_t e = { max_score, {0, 0, 0, 0}, pair_0, pair_1 };
            frontier_push(g, e);
            g->in_frontier[pair_0 * g->K + pair_1] = 1;
            continue;
        }
        if (pair_1 == start_i) {
            if (joints[pair_0].v > 0.0) continue;
            if (g->in_frontier[pair_1 * g->K + pair_0]) continue;
            frontier_entry_t e = { max_score, {0, 0, 0, 0}, pair_1, pair_0 };
            frontier_push(g, e);
            g->in_frontier[pair_1 * g->K + pair_0] = 1;
            continue;
        }
    }
}

typedef struct { const float* fwd; const int64_t* n_fwd; const float* bwd; const int64_t* n_bwd; int64_t cap; } caf_fb_t;

/* src/cifcaf.cpp:349-411 */
static joint_t connection_value(grow_ctx_t* g, const joint_t* joints, const caf_fb_t* fb,
                                int64_t start_i, int64_t end_i, int reverse_match_, double filter_sigmas) {
    int64_t caf_i = 0; int forward = 1;
    for (int64_t f = 0; f < g->C; f++) {
        int64_t pair_0 = g->skeleton[2 * f], pair_1 = g->skeleton[2 * f + 1];
        if (pair_0 == start_i && pair_1 == end_i) { forward = 1; break; }
        if (pair_1 == start_i && pair_0 == end_i) { forward = 0; break; }
        caf_i++;
    }
    const float* caf_f = (forward ? fb->fwd : fb->bwd) + caf_i * fb->cap * 7;
    int64_t n_f = (forward ? fb->n_fwd : fb->n_bwd)[caf_i];
    const float* caf_b = (forward ? fb->bwd : fb->fwd) + caf_i * fb->cap * 7;
    int64_t n_b = (forward ? fb->n_bwd : fb->n_fwd)[caf_i];
    int only_max = 0;

    const joint_t* start_j = &joints[start_i];
    joint_t new_j = grow_connection_blend(caf_f, n_f, start_j->x, start_j->y, start_j->s, filter_sigmas, only_max);
    if (new_j.v == 0.0) return new_j;

    new_j.v = sqrt(new_j.v * start_j->v);
    if (new_j.v < g->p->keypoint_threshold || new_j.v < start_j->v * g->p->keypoint_threshold_rel) {
        new_j.v = 0.0;
        return new_j;
    }
    if (g->p->reverse_match && reverse_match_ && start_i < g->occ_n_fields) {
        joint_t reverse_j = grow_connection_blend(caf_b, n_b, new_j.x, new_j.y, new_j.s, filter_sigmas, only_max);
        if (reverse_j.v == 0.0) { new_j.v = 0.0; return new_j; }
        if (fabs(start_j->x - reverse_j.x) + fabs(start_j->y - reverse_j.y) > start_j->s) {
            new_j.v = 0.0; return new_j;
        }
    }
    return new_j;
}

/* src/cifcaf.cpp:265-313 */
static void grow(grow_ctx_t* g, joint_t* joints, const caf_fb_t* fb, int reverse_match_, double filter_sigmas) {
    g->heap_n = 0;
    memset(g->in_frontier, 0, (size_t)(g->K * g->K));
    for (int64_t j = 0; j < g->K; j++) {
        if (joints[j].v == 0.0) continue;
        frontier_add_from(g, joints, j);
    }
    while (g->heap_n > 0) {
        frontier_entry_t entry = frontier_pop(g);
        if (joints[entry.end_i].v > 0.0) continue;
        if (entry.joint.v == 0.0) {
            joint_t new_joint = connection_value(g, joints, fb, entry.start_i, entry.end_i, reverse_match_, filter_sigmas);
            if (new_joint.v == 0.0) continue;   /* block_joints branch has no effect */
            if (!g->p->greedy) {
                frontier_entry_t e = { (float)new_joint.v, new_joint, entry.start_i, entry.end_i };
                frontier_push(g, e);
                continue;
            }
            entry.max_score = (float)new_joint.v;
            entry.joint = new_joint;
        }
        joints[entry.end_i] = entry.joint;
        frontier_add_from(g, joints, entry.end_i);
    }
}

/* src/cifcaf.cpp:429-449 */
static void flood_fill(grow_ctx_t* g, joint_t* joints) {
    g->heap_n = 0;
    memset(g->in_frontier, 0, (size_t)(g->K * g->K));
    for (int64_t j = 0; j < g->K; j++) {
        if (joints[j].v == 0.0) continue;
        frontier_add_from(g, joints, j);
    }
    while (g->heap_n > 0) {
        frontier_entry_t entry = frontier_pop(g);
        if (joints[entry.end_i].v > 0.0) continue;
        joints[entry.end_i] = joints[entry.start_i];
        joints[entry.end_i].v = 0.00001;
        frontier_add_from(g, joints, entry.end_i);
    }
}

/* include/.../nms_keypoints.hpp:25-32: running sum truncated to float each step */
static double uniform_score(const joint_t* joints, int64_t K) {
    double init = 0.0;
    for (int64_t k = 0; k < K; k++) { float i = (float)init; init = (double)i + joints[k].v; }
    return init / (double)K;
}

typedef struct { joint_t* joints; int64_t id; double score; int64_t order; } ann_t;

static int ann_cmp(const void* a, const void* b) {
    const ann_t* x = (const ann_t*)a; const ann_t* y = (const ann_t*)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return x->order < y->order ? -1 : (x->order > y->order ? 1 : 0);
}

/* src/nms_keypoints.cpp:17-69; returns new count */
static int64_t nms_keypoints(occupancy_t* occ, ann_t* anns, int64_t n, int64_t K, const oracle_params_t* p) {
    occupancy_clear(occ);
    for (int64_t a = 0; a < n; a++) { anns[a].score = uniform_score(anns[a].joints, K); anns[a].order = a; }
    qsort(anns, (size_t)n, sizeof(ann_t), ann_cmp);
    for (int64_t a = 0; a < n; a++) {
        for (int64_t f = 0; f < K; f++) {
            if (f >= occ->F) break;
            joint_t* j = &anns[a].joints[f];
            if (j->v == 0.0) continue;
            if (occupancy_get(occ, f, j->x, j->y)) j->v *= p->nms_suppression;
            else occupancy_set(occ, f, j->x, j->y, j->s);
        }
    }
    for (int64_t a = 0; a < n; a++)
        for (int64_t f = 0; f < K; f++)
            if (!(anns[a].joints[f].v > p->nms_keypoint_threshold)) anns[a].joints[f].v = 0.0;
    int64_t m = 0;
    for (int64_t a = 0; a < n; a++) {
        double s = uniform_score(anns[a].joints, K);
        if (s < p->nms_instance_threshold) continue;
        anns[m] = anns[a]; anns[m].score = s; anns[m].order = m; m++;
    }
    qsort(anns, (size_t)m, sizeof(ann_t), ann_cmp);
    return m;
}

/*
 * src/cifcaf.cpp:126-262 CifCaf::call_with_initial_annotations, fresh instance.
 *
 * cif [F][5][h][w], caf [C][8][h][w] (f32, contiguous); skeleton [C][2] 0-based.
 * initial_annotations [n_init][K][4] (v,x,y,s) or NULL; initial_ids [n_init].
 * Outputs: out_ann [cap][K][4] f32 (v,x,y,s), out_ids [cap]; returns N (may
 * exceed cap, in which case only cap are written).
 * Optional taps (NULL to skip): tap_cifhr [F][H][W]; tap_seeds_f/vxys (cap
 * tap_seeds_cap, count in *tap_n_seeds); tap_fwd/tap_bwd [C][h*w][7] with
 * counts tap_n_fwd/tap_n_bwd [C]; *tap_n_pre_nms = annotations before NMS.
 */
int64_t oracle_cifcaf_call(const float* cif, int64_t F, int64_t cif_h, int64_t cif_w, int64_t cif_stride,
                           const float* caf, int64_t C, int64_t caf_h, int64_t caf_w, int64_t caf_stride,
                           const int64_t* skeleton, int64_t n_keypoints,
                           const float* initial_annotations, const int64_t* initial_ids, int64_t n_init,
                           const oracle_params_t* p,
                           float* out_ann, int64_t* out_ids, int64_t cap,
                           float* tap_cifhr,
                           int64_t* tap_seeds_f, float* tap_seeds_vxys, int64_t tap_seeds_cap, int64_t* tap_n_seeds,
                           float* tap_fwd, int64_t* tap_n_fwd, float* tap_bwd, int64_t* tap_n_bwd,
                           int64_t* tap_n_pre_nms) {
    int64_t K = n_keypoints;
    int64_t H = (cif_h - 1) * cif_stride + 1, W = (cif_w - 1) * cif_stride + 1;

    /* cifhr.reset + accumulate: cifcaf.cpp:140-142 */
    float* acc = (float*)calloc((size_t)(F * H * W), sizeof(float));
    oracle_cifhr_accumulate(cif, F, cif_h, cif_w, cif_stride, 0.0, 1.0, p, acc);
    if (tap_cifhr) memcpy(tap_cifhr, acc, sizeof(float) * (size_t)(F * H * W));

    /* seeds: cifcaf.cpp:144-148 */
    int64_t seeds_cap = F * cif_h * cif_w;
    int64_t* seeds_f = (int64_t*)malloc(sizeof(int64_t) * (size_t)(seeds_cap + 1));
    float* seeds_vxys = (float*)malloc(sizeof(float) * 4 * (size_t)(seeds_cap + 1));
    int64_t n_seeds = oracle_cifseeds(cif, F, cif_h, cif_w, cif_stride, acc, p, seeds_f, seeds_vxys, seeds_cap);
    if (tap_n_seeds) *tap_n_seeds = n_seeds;
    if (tap_seeds_f) for (int64_t k = 0; k < n_seeds && k < tap_seeds_cap; k++) {
        tap_seeds_f[k] = seeds_f[k]; memcpy(tap_seeds_vxys + 4 * k, seeds_vxys + 4 * k, 16);
    }

    /* caf scored: cifcaf.cpp:153-161 */
    int64_t caf_hw = caf_h * caf_w;
    float* fwd = (float*)malloc(sizeof(float) * 7 * (size_t)(C * caf_hw + 1));
    float* bwd = (float*)malloc(sizeof(float) * 7 * (size_t)(C * caf_hw + 1));
    int64_t* n_fwd = (int64_t*)calloc((size_t)C + 1, sizeof(int64_t));
    int64_t* n_bwd = (int64_t*)calloc((size_t)C + 1, sizeof(int64_t));
    oracle_cafscored(caf, C, caf_h, caf_w, caf_stride, skeleton, acc, F, H, W, -1.0, p, fwd, n_fwd, bwd, n_bwd);
    if (tap_fwd) {
        memcpy(tap_fwd, fwd, sizeof(float) * 7 * (size_t)(C * caf_hw));
        memcpy(tap_bwd, bwd, sizeof(float) * 7 * (size_t)(C * caf_hw));
        memcpy(tap_n_fwd, n_fwd, sizeof(int64_t) * (size_t)C);
        memcpy(tap_n_bwd, n_bwd, sizeof(int64_t) * (size_t)C);
    }
    caf_fb_t fb = { fwd, n_fwd, bwd, n_bwd, caf_hw };

    occupancy_t occ;
    occupancy_reset(&occ, F, H, W, p);   /* cifcaf.cpp:173 */

    grow_ctx_t g;
    g.K = K; g.C = C; g.skeleton = skeleton; g.p = p; g.occ_n_fields = occ.F;
    g.heap = (frontier_entry_t*)malloc(sizeof(frontier_entry_t) * (size_t)(4 * C + 8));
    g.heap_n = 0;
    g.in_frontier = (uint8_t*)malloc((size_t)(K * K));

    int64_t ann_cap = n_seeds + n_init + 1;
    ann_t* anns = (ann_t*)malloc(sizeof(ann_t) * (size_t)ann_cap);
    joint_t* joints_pool = (joint_t*)calloc((size_t)(ann_cap * K), sizeof(joint_t));
    int64_t n_ann = 0;

    /* initial annotations: cifcaf.cpp:177-202 */
    for (int64_t a = 0; a < n_init; a++) {
        joint_t* joints = joints_pool + n_ann * K;
        for (int64_t k = 0; k < K; k++) {
            const float* s = initial_annotations + (a * K + k) * 4;
            joints[k].v = s[0]; joints[k].x = s[1]; joints[k].y = s[2]; joints[k].s = s[3];
        }
        grow(&g, joints, &fb, 1, 1.0);
        for (int64_t of = 0; of < occ.F; of++) {
            if (joints[of].v == 0.0) continue;
            occupancy_set(&occ, of, joints[of].x, joints[of].y, joints[of].s);
        }
        anns[n_ann].joints = joints; anns[n_ann].id = initial_ids[a]; n_ann++;
    }

    /* seed loop: cifcaf.cpp:204-231 */
    for (int64_t si = 0; si < n_seeds; si++) {
        int64_t f = seeds_f[si];
        float x = seeds_vxys[4 * si + 1], y = seeds_vxys[4 * si + 2], s = seeds_vxys[4 * si + 3];
        if (occupancy_get(&occ, f, x, y)) continue;
        joint_t* joints = joints_pool + n_ann * K;
        joints[f].v = seeds_vxys[4 * si + 0]; joints[f].x = x; joints[f].y = y; joints[f].s = s;
        grow(&g, joints, &fb, 1, 1.0);
        for (int64_t of = 0; of < occ.F; of++) {
            if (joints[of].v == 0.0) continue;
            occupancy_set(&occ, of, joints[of].x, joints[of].y, joints[of].s);
        }
        anns[n_ann].joints = joints; anns[n_ann].id = -1; n_ann++;
    }

    /* force complete: cifcaf.cpp:233-236, 414-426 */
    if (p->force_complete) {
        oracle_cafscored(caf, C, caf_h, caf_w, caf_stride, skeleton, acc, F, H, W,
                         p->force_complete_caf_th, p, fwd, n_fwd, bwd, n_bwd);
        for (int64_t a = 0; a < n_ann; a++) grow(&g, anns[a].joints, &fb, 0, 4.0);
        for (int64_t a = 0; a < n_ann; a++) flood_fill(&g, anns[a].joints);
    }
    if (tap_n_pre_nms) *tap_n_pre_nms = n_ann;

    int64_t keep = nms_keypoints(&occ, anns, n_ann, K, p);   /* cifcaf.cpp:241 */

    /* pack: cifcaf.cpp:246-261 */
    for (int64_t a = 0; a < keep && a < cap; a++) {
        for (int64_t k = 0; k < K; k++) {
            out_ann[(a * K + k) * 4 + 0] = (float)anns[a].joints[k].v;
            out_ann[(a * K + k) * 4 + 1] = (float)anns[a].joints[k].x;
            out_ann[(a * K + k) * 4 + 2] = (float)anns[a].joints[k].y;
            out_ann[(a * K + k) * 4 + 3] = (float)anns[a].joints[k].s;
        }
        out_ids[a] = anns[a].id;
    }

    free(joints_pool);
    free(g.heap); free(g.in_frontier);
    free(occ.occ); free(fwd); free(bwd); free(n_fwd); free(n_bwd);
    free(seeds_f); free(seeds_vxys); free(acc);
    free(anns);
    return keep;
}


/* ------------------------------------------------------------------ CifDet */

/* src/cif_hr.cpp:124-150 (CifDetHr::accumulate).  field is [F][6][h][w]: conf at 1, x,y at 2,3, w,h at 4,5
 * (headmeta.py:117-134: n_confidences 1, n_vectors 2, vector_offsets [True, False]). */
void oracle_cifdethr_accumulate(const float* field, int64_t F, int64_t h, int64_t w, int64_t stride,
                                double min_scale, double factor, const oracle_params_t* p, float* acc) {
    int64_t H = (h - 1) * stride + 1, W = (w - 1) * stride + 1;
    float min_scale_f = (float)(min_scale / (double)stride);
    int64_t hw = h * w;
    for (int64_t f = 0; f < F; f++) {
        const float* cf = field + f * 6 * hw;
        for (int64_t j = 0; j < h; j++) {
            for (int64_t i = 0; i < w; i++) {
                float v = cf[1 * hw + j * w + i];
                if ((double)v < p->cifhr_threshold) continue;
                float bw = cf[4 * hw + j * w + i];
                float bh = cf[5 * hw + j * w + i];
                if (bw < min_scale_f || bh < min_scale_f) continue;
                float x = cf[2 * hw + j * w + i] * (float)stride;
                float y = cf[3 * hw + j * w + i] * (float)stride;
                float sigma = fmaxf(1.0f, (float)(0.1 * (double)fminf(bw, bh) * (double)stride));
                float vn = (float)((double)(v / (float)p->cifhr_neighbors) * factor);
                cifhr_add_gauss(acc, H, W, p->cifhr_revision, f, vn, x, y, sigma, 1.0f);
            }
        }
    }
}

/* src/cifdet.cpp:24-80 on a FRESH instance (revision p->cifhr_revision), with CifDetSeeds::fill/get
 * (src/cif_seeds.cpp:69-90, 117-139).  p->seed_threshold plays CifDetSeeds::threshold.
 * Outputs: categories [cap] (f + 1), scores [cap], boxes [cap][4] (x1, y1, x2, y2); returns N <= max_detections.
 * Optional taps: tap_cifhr [F][H][W]; tap_seeds_f / tap_seeds_vxywh [.][5] (cap tap_seeds_cap), *tap_n_seeds. */
int64_t oracle_cifdet_call(const float* field, int64_t F, int64_t h, int64_t w, int64_t stride,
                           const oracle_params_t* p, int64_t max_detections_before_nms,
                           int64_t* out_categories, float* out_scores, float* out_boxes, int64_t cap,
                           float* tap_cifhr, int64_t* tap_seeds_f, float* tap_seeds_vxywh, int64_t tap_seeds_cap,
                           int64_t* tap_n_seeds) {
    int64_t H = (h - 1) * stride + 1, W = (w - 1) * stride + 1, hw = h * w;
    float* acc = (float*)calloc((size_t)(F * H * W), sizeof(float));
    oracle_cifdethr_accumulate(field, F, h, w, stride, 0.0, 1.0, p, acc);      /* cifdet.cpp:31 */
    if (tap_cifhr) memcpy(tap_cifhr, acc, sizeof(float) * (size_t)(F * H * W));

    seed_t* seeds = (seed_t*)malloc(sizeof(seed_t) * (size_t)(F * hw + 1));
    int64_t n = 0;
    for (int64_t f = 0; f < F; f++) {                                          /* cif_seeds.cpp:69-90 */
        const float* cf = field + f * 6 * hw;
        for (int64_t j = 0; j < h; j++) {
            for (int64_t i = 0; i < w; i++) {
                float c = cf[1 * hw + j * w + i];
                if ((double)c < p->seed_threshold) continue;
                float x = cf[2 * hw + j * w + i] * (float)stride;
                float y = cf[3 * hw + j * w + i] * (float)stride;
                float v = (float)(0.9 * (double)cifhr_value(acc, F, H, W, p->cifhr_revision, f, x, y, -1.0f)
                                  + 0.1 * (double)c);
                if ((double)v < p->seed_threshold) continue;
                seeds[n].f = f; seeds[n].v = v; seeds[n].x = x; seeds[n].y = y;
                seeds[n].s = cf[4 * hw + j * w + i] * (float)stride;
                seeds[n].h = cf[5 * hw + j * w + i] * (float)stride;
                seeds[n].order = n;
                n++;
            }
        }
    }
    if (p->seed_sort_stable) qsort(seeds, (size_t)n, sizeof(seed_t), seed_cmp);
    else seed_std_sort(seeds, n);                                              /* cif_seeds.cpp:117-121 */
    if (tap_n_seeds) *tap_n_seeds = n;
    if (tap_seeds_f) for (int64_t k = 0; k < n && k < tap_seeds_cap; k++) {
        tap_seeds_f[k] = seeds[k].f;
        tap_seeds_vxywh[5 * k + 0] = seeds[k].v; tap_seeds_vxywh[5 * k + 1] = seeds[k].x;
        tap_seeds_vxywh[5 * k + 2] = seeds[k].y; tap_seeds_vxywh[5 * k + 3] = seeds[k].s;
        tap_seeds_vxywh[5 * k + 4] = seeds[k].h;
    }

    occupancy_t occ;
    occupancy_reset(&occ, F, H, W, p);                                         /* cifdet.cpp:43 */
    int64_t n_det = 0;
    for (int64_t si = 0; si < n; si++) {                                       /* cifdet.cpp:50-66 */
        const int64_t f = seeds[si].f;
        const float c = seeds[si].v, x = seeds[si].x, y = seeds[si].y, bw = seeds[si].s, bh = seeds[si].h;
        if (occupancy_get(&occ, f, x, y)) continue;
        occupancy_set(&occ, f, x, y, 0.1 * (double)fminf(bw, bh));
        if (n_det < cap) {
            out_categories[n_det] = f + 1;
            out_scores[n_det] = c;
            out_boxes[4 * n_det + 0] = x - 0.5f * bw; out_boxes[4 * n_det + 1] = y - 0.5f * bh;
            out_boxes[4 * n_det + 2] = x + 0.5f * bw; out_boxes[4 * n_det + 3] = y + 0.5f * bh;
        }
        n_det++;
        if (n_det >= max_detections_before_nms) break;
    }
    free(occ.occ); free(seeds); free(acc);
    return n_det;
}
