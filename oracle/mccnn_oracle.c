/*
 * mccnn_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * CPU restatement, in plain C, of the stereo-matching hot path of
 * Jackie-Chou/MC-CNN-python (src/process_functional.py, src/model.py).  It is the
 * checker that the CUDA path is compared against.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Parity pinning: the reference ships no tests or golden vectors.
 * This restatement is pinned against the reference's OWN NumPy code executed in the
 * authoring container (oracle/ref_loader.py execs /root/reference/src/process_functional.py
 * under NumPy 2.3.5); the resulting vectors are committed under tests/golden/ (generator:
 * oracle/gen_golden.py) and re-checked against this file on every test run.  Every stage
 * after the CNN reproduces the reference BIT-EXACTLY, including NumPy's float32 pairwise
 * summation order; the CNN (TensorFlow is not installable here) is restated from
 * model.py and pinned against a torch-CPU conv2d restatement -- "parity unpinned" for the
 * CNN with respect to TensorFlow itself.
 *
 * All arrays are float32, C-contiguous, laid out exactly as the reference lays them out:
 *   volumes [D][H][W], images [H][W], features [H][W][C], disparity maps [H][W].
 * "pf:N" cites /root/reference/src/process_functional.py line N.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).  FP contraction must stay
 * off: the bit-exact stages rely on separately rounded float32 operations.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX3(d, h, w) (((size_t)(d) * H + (size_t)(h)) * W + (size_t)(w))

int mccnn_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void mccnn_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* NumPy's float32 pairwise summation for a contiguous run (numpy/_core/src/umath/
 * loops_utils.h, FLOAT_pairwise_sum), as used by np.sum on the 64-channel products (pf:89)
 * and on the 5x5 bilateral patches (pf:463, pf:466).  Verified against NumPy 2.3.5. */
static float np_pairwise_sum_f32(const float *a, int n) {
    if (n < 8) {
        float res = -0.0f;
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8];
        int i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum_f32(a, n2) + np_pairwise_sum_f32(a + n2, n - n2);
    }
}

/* ------------------------------------------------------------------------------------------
 * a1/a2  NET forward + compute_features   (model.py:40-64, :90-125; pf:15-73)
 * img [H][W] (already normalised); weights HWIO: w1 [3][3][1][F], w2..w5 [3][3][F][F]; biases [F].
 * Zero-pad (patch-1)/2 = nl pixels per side once (pf:20-25), then nl VALID 3x3 cross-correlations,
 * ReLU after all but the last (model.py:51-60), then x * rsqrt(max(sum x^2, 1e-12)) (model.py:64).
 * Accumulation is in double (neutral high-precision reference; TF's own order is unknowable).
 * ------------------------------------------------------------------------------------------ */
int mccnn_oracle_features(const float *img, int H, int W, int nl, int F,
                          const float *const *weights, const float *const *biases, float *out) {
    int pad = nl;
    int h = H + 2 * pad, w = W + 2 * pad, c = 1;
    float *cur = (float *)calloc((size_t)h * w, sizeof(float));
    if (!cur) return -1;
    for (int y = 0; y < H; y++)
        memcpy(cur + (size_t)(y + pad) * w + pad, img + (size_t)y * W, sizeof(float) * W);
    for (int l = 0; l < nl; l++) {
        int oh = h - 2, ow = w - 2;
        float *nxt = (float *)malloc((size_t)oh * ow * F * sizeof(float));
        if (!nxt) { free(cur); return -1; }
        const float *wt = weights[l];
        const float *bs = biases[l];
#pragma omp parallel for schedule(static)
        for (int y = 0; y < oh; y++) {
            double *acc = (double *)malloc(sizeof(double) * F);
            for (int x = 0; x < ow; x++) {
                for (int o = 0; o < F; o++) acc[o] = 0.0;
                for (int ky = 0; ky < 3; ky++)
                    for (int kx = 0; kx < 3; kx++) {
                        const float *ip = cur + ((size_t)(y + ky) * w + (x + kx)) * c;
                        const float *wp = wt + (size_t)((ky * 3 + kx) * c) * F;
                        for (int i = 0; i < c; i++) {
                            double v = ip[i];
                            const float *wr = wp + (size_t)i * F;
                            for (int o = 0; o < F; o++) acc[o] += v * (double)wr[o];
                        }
                    }
                float *op = nxt + ((size_t)y * ow + x) * F;
                for (int o = 0; o < F; o++) {
                    float v = (float)(acc[o] + (double)bs[o]);
                    if (l < nl - 1 && v < 0.0f) v = 0.0f;
                    op[o] = v;
                }
            }
            free(acc);
        }
        free(cur);
        cur = nxt; h = oh; w = ow; c = F;
    }
    /* l2_normalize over channels, epsilon 1e-12 under max (model.py:64) */
#pragma omp parallel for schedule(static)
    for (int p = 0; p < H * W; p++) {
        const float *ip = cur + (size_t)p * F;
        double ss = 0.0;
        for (int o = 0; o < F; o++) ss += (double)ip[o] * (double)ip[o];
        double inv = 1.0 / sqrt(ss > 1e-12 ? ss : 1e-12);
        for (int o = 0; o < F; o++) out[(size_t)p * F + o] = (float)((double)ip[o] * inv);
    }
    free(cur);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a3  compute_cost_volume   (pf:78-113)
 * ------------------------------------------------------------------------------------------ */
int mccnn_oracle_cost_volume(const float *fl, const float *fr, int H, int W, int C, int D,
                             float *L, float *R) {
    if (D < 1 || W < D + 2) return -1;
    memset(L, 0, sizeof(float) * (size_t)D * H * W);
    memset(R, 0, sizeof(float) * (size_t)D * H * W);
    /* pf:87-91: score[d,h,w] = sum_c fl[h,w,c]*fr[h,w-d,c] for w >= d  (product array, then np.sum) */
#pragma omp parallel for collapse(2) schedule(static)
    for (int d = 0; d < D; d++)
        for (int h = 0; h < H; h++) {
            float *prod = (float *)malloc(sizeof(float) * C);
            for (int w = d; w < W; w++) {
                const float *a = fl + ((size_t)h * W + w) * C;
                const float *b = fr + ((size_t)h * W + (w - d)) * C;
                for (int c = 0; c < C; c++) prod[c] = a[c] * b[c];
                L[IDX3(d, h, w)] = np_pairwise_sum_f32(prod, C);
            }
            free(prod);
        }
    /* pf:94-95: fill column d-1 of every plane dd >= d with mean of columns d..d+2, d descending */
#pragma omp parallel for collapse(2) schedule(static)
    for (int dd = 1; dd < D; dd++)
        for (int h = 0; h < H; h++)
            for (int c = dd - 1; c >= 0; c--) {
                float *row = L + IDX3(dd, h, 0);
                row[c] = ((row[c + 1] + row[c + 2]) + row[c + 3]) / 3.0f;
            }
    /* pf:103-106 */
#pragma omp parallel for collapse(2) schedule(static)
    for (int d = 0; d < D; d++)
        for (int h = 0; h < H; h++) {
            float *rrow = R + IDX3(d, h, 0);
            const float *lrow = L + IDX3(d, h, 0);
            for (int w = 0; w < W - d; w++) rrow[w] = lrow[w + d];
            for (int c = W - d; c < W; c++) rrow[c] = ((rrow[c - 3] + rrow[c - 2]) + rrow[c - 1]) / 3.0f;
        }
    /* pf:111-112 */
    size_t n = (size_t)D * H * W;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { L[i] = -1.0f * L[i]; R[i] = -1.0f * R[i]; }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a4  compute_cross_region  (pf:571-657), arm-length form.
 * arms[(h*W+w)*4 + {0,1,2,3}] = {up, down, left, right} (number of pixels beyond the anchor),
 * count[h*W+w] = |U(h,w)| = sum over the vertical arm of (left+right+1) of each spine pixel.
 * ------------------------------------------------------------------------------------------ */
int mccnn_oracle_cross_arms(const float *img, int H, int W, float tau, int dist,
                            uint8_t *arms, int32_t *count) {
    if (dist < 1 || dist > 255) return -1;
    if (!(tau > 0.0f)) return -2;   /* the anchor must pass its own test (asserts at pf:601, :628) */
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++)
        for (int w = 0; w < W; w++) {
            float cur = img[(size_t)h * W + w];
            int up = 0, down = 0, left = 0, right = 0, lim;
            /* sqrt(x*x) == |x| in IEEE arithmetic: np.linalg.norm of a 1-vector (pf:588) */
            lim = (dist < h + 1 ? dist : h + 1);                      /* pf:585 */
            for (int b = 1; b < lim; b++) { if (fabsf(cur - img[(size_t)(h - b) * W + w]) >= tau) break; up = b; }
            lim = (dist < H - h ? dist : H - h);                      /* pf:593 */
            for (int b = 1; b < lim; b++) { if (fabsf(cur - img[(size_t)(h + b) * W + w]) >= tau) break; down = b; }
            lim = (dist < w + 1 ? dist : w + 1);                      /* pf:612 */
            for (int b = 1; b < lim; b++) { if (fabsf(cur - img[(size_t)h * W + w - b]) >= tau) break; left = b; }
            lim = (dist < W - w ? dist : W - w);                      /* pf:620 */
            for (int b = 1; b < lim; b++) { if (fabsf(cur - img[(size_t)h * W + w + b]) >= tau) break; right = b; }
            uint8_t *a = arms + ((size_t)h * W + w) * 4;
            a[0] = (uint8_t)up; a[1] = (uint8_t)down; a[2] = (uint8_t)left; a[3] = (uint8_t)right;
        }
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++)
        for (int w = 0; w < W; w++) {
            const uint8_t *a = arms + ((size_t)h * W + w) * 4;
            int n = 0;
            for (int hh = h - a[0]; hh <= h + a[1]; hh++) {
                const uint8_t *s = arms + ((size_t)hh * W + w) * 4;
                n += s[2] + s[3] + 1;
            }
            count[(size_t)h * W + w] = n;
        }
    return 0;
}

/* Explicit region list in the reference's own enumeration order (pf:640-655), for the
 * compatibility view and for pinning the arm representation against the reference output.
 * region [H][W][max_num][2] int32 padded with -1, max_num = (2*dist)^2. */
int mccnn_oracle_cross_region_list(const uint8_t *arms, int H, int W, int dist, int32_t *region) {
    size_t max_num = (size_t)(2 * dist) * (2 * dist);
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++)
        for (int w = 0; w < W; w++) {
            int32_t *out = region + ((size_t)h * W + w) * max_num * 2;
            const uint8_t *a = arms + ((size_t)h * W + w) * 4;
            size_t n = 0;
            for (int k = 0; k <= a[0] + a[1]; k++) {
                int hh = (k <= a[0]) ? h - k : h + (k - a[0]);
                const uint8_t *s = arms + ((size_t)hh * W + w) * 4;
                for (int j = 0; j <= s[2] + s[3]; j++) {
                    int ww = (j <= s[2]) ? w - j : w + (j - s[2]);
                    out[2 * n] = hh; out[2 * n + 1] = ww; n++;
                }
            }
            for (; n < max_num; n++) { out[2 * n] = -1; out[2 * n + 1] = -1; }
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a5  cost_volume_aggregation, one volume  (pf:149-163 / :166-180)
 * Flat float32 running sum in the reference's enumeration order, then / count.
 * vol is left untouched; out receives the result after `iters` rounds.
 * ------------------------------------------------------------------------------------------ */
int mccnn_oracle_cbca(const float *vol, float *out, const uint8_t *arms, const int32_t *count,
                      int D, int H, int W, int iters) {
    size_t n = (size_t)D * H * W;
    if (iters <= 0) { memcpy(out, vol, n * sizeof(float)); return 0; }
    float *bufa = (float *)malloc(n * sizeof(float));
    float *bufb = (float *)malloc(n * sizeof(float));
    if (!bufa || !bufb) { free(bufa); free(bufb); return -1; }
    memcpy(bufa, vol, n * sizeof(float));
    for (int it = 0; it < iters; it++) {
        const float *src = bufa;
        float *dst = (it == iters - 1) ? out : bufb;
#pragma omp parallel for collapse(2) schedule(static)
        for (int d = 0; d < D; d++)
            for (int h = 0; h < H; h++)
                for (int w = 0; w < W; w++) {
                    const uint8_t *a = arms + ((size_t)h * W + w) * 4;
                    float sum = 0.0f;                                         /* pf:157 */
                    for (int k = 0; k <= a[0] + a[1]; k++) {
                        int hh = (k <= a[0]) ? h - k : h + (k - a[0]);
                        const uint8_t *s = arms + ((size_t)hh * W + w) * 4;
                        const float *row = src + IDX3(d, hh, 0);
                        for (int j = 0; j <= s[2] + s[3]; j++) {
                            int ww = (j <= s[2]) ? w - j : w + (j - s[2]);
                            sum += row[ww];                                   /* pf:160 */
                        }
                    }
                    /* pf:161: float32 / int32 -> computed in f64 then stored f32; identical to an
                     * f32 division (innocuous double rounding, 53 >= 2*24+2). */
                    dst[IDX3(d, h, w)] = (float)((double)sum / (double)count[(size_t)h * W + w]);
                }
        float *t = bufa; bufa = bufb; bufb = t;
    }
    free(bufa); free(bufb);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a6  semi_global_matching, one in-place pass  (pf:476-568)
 * P1, P2 arrive as doubles (the caller may have formed P1/V in f64, pf:204) and are rounded to
 * float32 here exactly like `sgm_P1*np.ones(float32)` does (pf:504-505).
 * ------------------------------------------------------------------------------------------ */
int mccnn_oracle_sgm_pass(float *vol, const float *img_left, const float *img_right,
                          int D, int H, int W, int rh, int rw,
                          double sgm_P1, double sgm_P2, double sgm_Q1, double sgm_Q2, double sgm_D,
                          int is_left) {
    if (rh * rw != 0 || D < 2) return -1;
    if (!((rh == 0 && (rw == 1 || rw == -1)) || (rw == 0 && (rh == 1 || rh == -1)))) return -1;
    int starth, endh, steph, startw, endw, stepw;
    if (rh >= 0) { starth = rh; endh = H; steph = 1; } else { starth = H + rh - 1; endh = -1; steph = -1; }
    if (rw >= 0) { startw = rw; endw = W; stepw = 1; } else { startw = W + rw - 1; endw = -1; stepw = -1; }
    const float P1a = (float)sgm_P1, P2a = (float)sgm_P2;
    const float Q1 = (float)sgm_Q1, Q2 = (float)sgm_Q2, tD = (float)sgm_D;
    const float P1q1 = P1a / Q1, P2q1 = P2a / Q1, P1q2 = P1a / Q2, P2q2 = P2a / Q2;
    const float *own = is_left ? img_left : img_right;
    const float *oth = is_left ? img_right : img_left;
    /* Scanlines along the non-moving axis are independent (each cell reads only its predecessor
     * on the same scanline); the visiting order of pf:545-546 is preserved along each scanline,
     * which is all the recurrence sees.  Threads therefore split scanlines. */
    const int horizontal = (rh == 0);
    const int nlines = horizontal ? H : W;
    int fail = 0;
#pragma omp parallel
    {
        float *p1 = (float *)malloc(sizeof(float) * D);
        float *p2 = (float *)malloc(sizeof(float) * D);
        float *prev = (float *)malloc(sizeof(float) * D);
        if (!p1 || !p2 || !prev) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for schedule(static)
        for (int line = 0; line < nlines; line++) {
            int h = horizontal ? line : starth, w = horizontal ? startw : line;
            int nsteps = horizontal ? (W - 1) : (H - 1);
            for (int s = 0; s < nsteps; s++, h += (horizontal ? 0 : steph), w += (horizontal ? stepw : 0)) {
            int qh = h - rh, qw = w - rw;
            float D1 = fabsf(own[(size_t)h * W + w] - own[(size_t)qh * W + qw]);   /* pf:512/525 */
            for (int d = 0; d < D; d++) {
                float D2 = 0.0f;
                if (is_left) {
                    if (!(w - d < 0 || w - rw - d < 0))                            /* pf:517 */
                        D2 = fabsf(oth[(size_t)h * W + (w - d)] - oth[(size_t)qh * W + (w - rw - d)]);
                } else {
                    if (!(w + d >= W || w - rw + d >= W))                          /* pf:530 */
                        D2 = fabsf(oth[(size_t)h * W + (w + d)] - oth[(size_t)qh * W + (w - rw + d)]);
                }
                int c1 = (D1 < tD) && (D2 < tD);
                int c2 = (D1 >= tD) && (D2 >= tD);
                if (c1) { p1[d] = P1a; p2[d] = P2a; }
                else if (c2) { p1[d] = P1q2; p2[d] = P2q2; }
                else { p1[d] = P1q1; p2[d] = P2q1; }
            }
            float m = INFINITY;
            for (int d = 0; d < D; d++) { prev[d] = vol[IDX3(d, qh, qw)]; if (prev[d] < m) m = prev[d]; }
            for (int d = 0; d < D; d++) {
                float item1 = prev[d];
                float item4 = m + p2[d];
                float best;
                if (d == 0) {
                    float item3 = prev[d + 1] + p1[d];
                    float t = (item4 < item3) ? item4 : item3;
                    best = (t < item1) ? t : item1;                                 /* pf:552 */
                } else if (d == D - 1) {
                    float item2 = prev[d - 1] + p1[d];
                    float t = (item2 < item1) ? item2 : item1;
                    best = (item4 < t) ? item4 : t;                                 /* pf:566 */
                } else {
                    float item2 = prev[d - 1] + p1[d];
                    float item3 = prev[d + 1] + p1[d];
                    float t1 = (item2 < item1) ? item2 : item1;
                    float t2 = (item4 < item3) ? item4 : item3;
                    best = (t2 < t1) ? t2 : t1;                                     /* pf:559 */
                }
                float *cell = vol + IDX3(d, h, w);
                *cell = (*cell + best) - m;
            }
            }
        }
        }
        free(p1); free(p2); free(prev);
    }
    (void)endh; (void)endw;
    return fail ? -1 : 0;
}

/* a7  SGM_average for one volume: four chained in-place passes (pf:194-210); the final
 * (X+X+X+X)/4. is the identity on finite float32 but is evaluated anyway. */
int mccnn_oracle_sgm_average(float *vol, const float *img_left, const float *img_right,
                             int D, int H, int W, double P1, double P2, double Q1, double Q2,
                             double tD, double V, int is_left) {
    int rc = 0;
    rc |= mccnn_oracle_sgm_pass(vol, img_left, img_right, D, H, W, 0, 1, P1, P2, Q1, Q2, tD, is_left);
    rc |= mccnn_oracle_sgm_pass(vol, img_left, img_right, D, H, W, 0, -1, P1, P2, Q1, Q2, tD, is_left);
    rc |= mccnn_oracle_sgm_pass(vol, img_left, img_right, D, H, W, -1, 0, P1 / V, P2, Q1, Q2, tD, is_left);
    rc |= mccnn_oracle_sgm_pass(vol, img_left, img_right, D, H, W, 1, 0, P1 / V, P2, Q1, Q2, tD, is_left);
    size_t n = (size_t)D * H * W;
    for (size_t i = 0; i < n; i++) { float x = vol[i]; vol[i] = (((x + x) + x) + x) / 4.0f; }
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * a8  disparity_prediction, one volume  (pf:245-254)
 * ------------------------------------------------------------------------------------------ */
int mccnn_oracle_wta(const float *vol, int D, int H, int W, float *disp) {
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int p = 0; p < H * W; p++) {
        float mc = INFINITY; int md = -1;
        for (int d = 0; d < D; d++) { float v = vol[(size_t)d * H * W + p]; if (v < mc) { mc = v; md = d; } }
        if (md < 0) bad = 1;                                                      /* pf:253 assert */
        disp[p] = (float)md;
    }
    return bad ? -1 : 0;
}

/* np.median of a small float32 list: NaN if any NaN; odd -> middle; even -> (a+b)/2 in float32. */
static float np_median_f32(float *v, int n) {
    for (int i = 0; i < n; i++) if (isnan(v[i])) return NAN;
    for (int i = 1; i < n; i++) { float x = v[i]; int j = i - 1; while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; j--; } v[j + 1] = x; }
    if (n & 1) return v[n / 2];
    return (v[n / 2 - 1] + v[n / 2]) / 2.0f;
}

/* ------------------------------------------------------------------------------------------
 * a9  interpolation  (pf:279-378).  labels (optional, may be NULL) receives the consistency map.
 * ------------------------------------------------------------------------------------------ */
int mccnn_oracle_interpolation(const float *dl, const float *dr, int H, int W, int ndisp,
                               float *out, int32_t *labels_out) {
    int32_t *lab = (int32_t *)calloc((size_t)H * W, sizeof(int32_t));
    if (!lab) return -1;
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++)
        for (int w = 0; w < W; w++) {
            int ld = (int)dl[(size_t)h * W + w];                                   /* pf:287 */
            if (w < ld) { lab[(size_t)h * W + w] = 2; continue; }
            float rd = dr[(size_t)h * W + (w - ld)];
            if (fabsf((float)ld - rd) <= 1.0f) continue;                           /* pf:294 */
            int lim = (w + 1 < ndisp) ? w + 1 : ndisp;
            int found = 0;
            for (int d = 0; d < lim; d++)
                if (fabsf((float)d - dr[(size_t)h * W + (w - d)]) <= 1.0f) { found = 1; break; }
            lab[(size_t)h * W + w] = found ? 1 : 2;
        }
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++)
        for (int w = 0; w < W; w++) {
            size_t p = (size_t)h * W + w;
            if (lab[p] == 0) { out[p] = dl[p]; continue; }
            if (lab[p] == 1) {
                float nb[4]; int cnt = 0;
                for (int w_ = w + 1; w_ < W; w_++) if (lab[(size_t)h * W + w_] == 0) { nb[cnt++] = dl[(size_t)h * W + w_]; break; }
                for (int w_ = w - 1; w_ >= 0; w_--) if (lab[(size_t)h * W + w_] == 0) { nb[cnt++] = dl[(size_t)h * W + w_]; break; }
                for (int h_ = h + 1; h_ < H; h_++) if (lab[(size_t)h_ * W + w] == 0) { nb[cnt++] = dl[(size_t)h_ * W + w]; break; }
                for (int h_ = h - 1; h_ >= 0; h_--) if (lab[(size_t)h_ * W + w] == 0) { nb[cnt++] = dl[(size_t)h_ * W + w]; break; }
                out[p] = cnt ? np_median_f32(nb, cnt) : dl[p];                    /* pf:353-356 */
            } else {
                float v = dl[p];
                for (int w_ = w + 1; w_ < W; w_++) if (lab[(size_t)h * W + w_] == 0) { v = dl[(size_t)h * W + w_]; break; }
                out[p] = v;                                                       /* pf:365-373 */
            }
        }
    if (labels_out) memcpy(labels_out, lab, sizeof(int32_t) * (size_t)H * W);
    free(lab);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a10  subpixel_enhance  (pf:381-400).  NumPy >= 2 scalar rules: every operation is float32.
 * ------------------------------------------------------------------------------------------ */
int mccnn_oracle_subpixel(const float *disp, const float *vol, int D, int H, int W, float *out) {
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++)
        for (int w = 0; w < W; w++) {
            float d = disp[(size_t)h * W + w];
            int im = (int)(d - 1.0f), ip = (int)(d + 1.0f), ic = (int)d;          /* int() truncates */
            if (im < 0 || ip >= D) { out[(size_t)h * W + w] = d; continue; }
            float Cm = vol[IDX3(im, h, w)], Cp = vol[IDX3(ip, h, w)], C = vol[IDX3(ic, h, w)];
            float num = Cp - Cm;
            float den = 2.0f * ((Cp - 2.0f * C) + Cm);
            out[(size_t)h * W + w] = d - num / den;                                /* pf:396 */
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a11  median_filter  (pf:403-421): border-clipped window, np.median.
 * ------------------------------------------------------------------------------------------ */
int mccnn_oracle_median(const float *in, int H, int W, int fh, int fw, float *out) {
    int rh = (fh - 1) / 2, rw = (fw - 1) / 2;
    if (fh * fw > 1024) return -1;
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++)
        for (int w = 0; w < W; w++) {
            float buf[1024];
            int hs = h - rh < 0 ? 0 : h - rh, he = h + rh + 1 > H ? H : h + rh + 1;
            int ws = w - rw < 0 ? 0 : w - rw, we = w + rw + 1 > W ? W : w + rw + 1;
            int n = 0;
            for (int y = hs; y < he; y++) for (int x = ws; x < we; x++) buf[n++] = in[(size_t)y * W + x];
            out[(size_t)h * W + w] = np_median_f32(buf, n);
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a12  bilateral_filter  (pf:424-470).  `table` [fh][fw] is the float32 weight table that the
 * reference builds in float64 and stores as float32 (pf:433-436); the caller builds it the same
 * way (oracle.py: bilateral_table).  Sums follow NumPy's pairwise order over the clipped patch.
 * ------------------------------------------------------------------------------------------ */
int mccnn_oracle_bilateral(const float *img, const float *in, int H, int W, int fh, int fw,
                           const float *table, float blur_threshold, float *out) {
    int ch = (fh - 1) / 2, cw = (fw - 1) / 2;
    if (fh * fw > 1024) return -1;
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++)
        for (int w = 0; w < W; w++) {
            float wts[1024], prod[1024];
            int hs = h - ch < 0 ? 0 : h - ch, he = h + ch + 1 > H ? H : h + ch + 1;
            int ws = w - cw < 0 ? 0 : w - cw, we = w + cw + 1 > W ? W : w + cw + 1;
            float cur = img[(size_t)h * W + w];
            int n = 0;
            for (int y = hs; y < he; y++)
                for (int x = ws; x < we; x++) {
                    float diff = fabsf(img[(size_t)y * W + x] - cur);               /* pf:458-459 */
                    float mask = (diff < blur_threshold) ? 1.0f : 0.0f;            /* pf:460 */
                    float wt = mask * table[(ch - (h - y)) * fw + (cw - (w - x))];  /* pf:449-453,462 */
                    wts[n] = wt;
                    prod[n] = wt * in[(size_t)y * W + x];                          /* pf:465 */
                    n++;
                }
            float wsum = np_pairwise_sum_f32(wts, n);
            out[(size_t)h * W + w] = np_pairwise_sum_f32(prod, n) / wsum;          /* pf:466 */
        }
    return 0;
}
