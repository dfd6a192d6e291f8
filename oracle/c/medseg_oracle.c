/* medseg_oracle.c -- ORACLE (test infrastructure only; never linked into the product).
 *
 * Plain-C, OpenCV-free restatement of the integer / byte stages of the reference's per-slice path.
 * Paths are relative to /root/reference.  The arithmetic the reference delegates to OpenCV
 * (un-vendored, version unpinned) is restated from the published algorithms:
 *   - connected components: flood fill (any correct labelling gives the same masks; the reference's
 *     results do not depend on label numbering),
 *   - morphologyEx(OPEN, 3x3 rect) with OpenCV's default border rule (outside pixels never win),
 *   - findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE): Suzuki-Abe border following, closed form
 *     in SURVEY.md section 8(c).
 * PARITY PINNING: the reference ships no golden vectors; this file is pinned against cv2 4.13.0
 * (tests/test_oracle_c.py, thousands of random + adversarial masks) and against tests/golden/.
 *
 * Build: oracle/c/Makefile  ->  oracle/_build/libmedseg_oracle.so   (gcc -O2 -ffp-contract=off)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ------------------------------------------------------------------ preprocess (src/preprocess.cpp) */

/* src/preprocess.cpp:65-74 */
static void compute_minmax(const uint16_t* src, size_t len, uint16_t* mn, uint16_t* mx) {
    uint16_t a = 0xFFFF, b = 0;
    for (size_t i = 0; i < len; ++i) {
        uint16_t v = src[i];
        if (v < a) a = v;
        if (v > b) b = v;
    }
    *mn = a;
    *mx = b;
}

/* src/preprocess.cpp:81-118 (outW/outH are 512 in the reference, :81) */
void orc_preprocess(const uint16_t* src, int w, int h, int outW, int outH, uint8_t* dst) {
    const double stepX = (double)w / outW;                       /* :82 */
    const double stepY = (double)h / outH;                       /* :83 */
    uint16_t mn, mx;
    compute_minmax(src, (size_t)w * h, &mn, &mx);                /* :91 */
    if (mn == mx) mx = mn + 1;                                   /* :92 */
    const double scale8 = 255.0 / (mx - mn);                     /* :93 */
    for (int y = 0; y < outH; ++y) {
        for (int x = 0; x < outW; ++x) {
            double fx = x * stepX, fy = y * stepY;               /* :100 */
            int ix = (int)fx;
            int iy = (int)fy;
            int ix1 = ix + 1 < w - 1 ? ix + 1 : w - 1;           /* :103 */
            int iy1 = iy + 1 < h - 1 ? iy + 1 : h - 1;           /* :104 */
            double dx = fx - ix, dy = fy - iy;
            uint16_t v00 = src[(size_t)iy * w + ix];
            uint16_t v01 = src[(size_t)iy * w + ix1];
            uint16_t v10 = src[(size_t)iy1 * w + ix];
            uint16_t v11 = src[(size_t)iy1 * w + ix1];
            double v = (1 - dx) * (1 - dy) * v00 + dx * (1 - dy) * v01 + (1 - dx) * dy * v10 + dx * dy * v11; /* :112-115 */
            dst[(size_t)y * outW + x] = (uint8_t)((v - mn) * scale8 + 0.5);  /* :116 */
        }
    }
}

/* src/process.cpp:36-39 */
void orc_u8_to_float(const uint8_t* src, size_t n, float* dst) {
    for (size_t i = 0; i < n; ++i) dst[i] = (float)src[i] / 255.0f;
}

/* src/process.cpp:158-170: strict > from -FLT_MAX over the first `nc` planes of `logits` [C][n] */
void orc_argmax(const float* logits, int nc, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; ++i) {
        float best = -FLT_MAX;
        uint8_t idx = 0;
        for (int c = 0; c < nc; ++c) {
            float v = logits[(size_t)c * n + i];
            if (v > best) {
                best = v;
                idx = (uint8_t)c;
            }
        }
        out[i] = idx;
    }
}

/* src/process.cpp:178-185 */
void orc_mask_to_image(const uint8_t* mask, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; ++i) out[i] = mask[i] == 1 ? 128 : (mask[i] == 2 ? 255 : 0);
}

/* ------------------------------------------------------------------ connected components */

static const int DX8[8] = {1, 1, 0, -1, -1, -1, 0, 1};   /* 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE */
static const int DY8[8] = {0, -1, -1, -1, 0, 1, 1, 1};

/* Labels nonzero pixels of `bin` (conn = 4 or 8) in raster order of first pixel, labels from 1;
 * 0 = background.  Returns the number of components.  stats[i] = {left, top, right, bottom, area}
 * for label i (stats may be NULL). */
int orc_ccl(const uint8_t* bin, int H, int W, int conn, int32_t* labels, int32_t* stats, int stats_cap) {
    size_t n = (size_t)H * W;
    memset(labels, 0, n * sizeof(int32_t));
    int32_t* stack = (int32_t*)malloc(n * sizeof(int32_t));
    int nc = 0;
    for (size_t s = 0; s < n; ++s) {
        if (!bin[s] || labels[s]) continue;
        ++nc;
        int left = W, top = H, right = -1, bottom = -1, area = 0;
        size_t sp = 0;
        stack[sp++] = (int32_t)s;
        labels[s] = nc;
        while (sp) {
            int p = stack[--sp];
            int x = p % W, y = p / W;
            ++area;
            if (x < left) left = x;
            if (x > right) right = x;
            if (y < top) top = y;
            if (y > bottom) bottom = y;
            for (int d = 0; d < 8; ++d) {
                if (conn == 4 && (d & 1)) continue;
                int xx = x + DX8[d], yy = y + DY8[d];
                if (xx < 0 || yy < 0 || xx >= W || yy >= H) continue;
                int q = yy * W + xx;
                if (bin[q] && !labels[q]) {
                    labels[q] = nc;
                    stack[sp++] = q;
                }
            }
        }
        if (stats && nc < stats_cap) {
            int32_t* st = stats + (size_t)nc * 5;
            st[0] = left; st[1] = top; st[2] = right; st[3] = bottom; st[4] = area;
        }
    }
    free(stack);
    return nc;
}

/* ------------------------------------------------------------------ postprocess (src/postprocess.cpp) */

static void morph3(const uint8_t* src, uint8_t* dst, int H, int W, int erode) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            uint8_t v = erode ? 255 : 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    int xx = x + dx, yy = y + dy;
                    if (xx < 0 || yy < 0 || xx >= W || yy >= H) continue; /* default border never wins */
                    uint8_t s = src[(size_t)yy * W + xx];
                    if (erode ? s < v : s > v) v = s;
                }
            dst[(size_t)y * W + x] = v;
        }
}

/* src/postprocess.cpp:47-79 (with fill_holes_inside_foreground :13-44 inlined) */
void orc_postprocess(const uint8_t* src, uint8_t* out, int H, int W, int fg, float ratio) {
    size_t n = (size_t)H * W;
    uint8_t* mask = (uint8_t*)malloc(n);
    uint8_t* bin = (uint8_t*)malloc(n);
    uint8_t* tmp = (uint8_t*)malloc(n);
    int32_t* labels = (int32_t*)malloc(n * sizeof(int32_t));
    int cap = (int)(n / 1) + 2;
    int32_t* stats = (int32_t*)malloc((size_t)cap * 5 * sizeof(int32_t));
    memcpy(mask, src, n);
    const int min_area = (int)(W * H * ratio);                        /* :30 / :66 (int*int -> float multiply) */

    for (size_t i = 0; i < n; ++i) bin[i] = mask[i] == fg ? 0 : 255;   /* :18-22 inv = ~(mask == FG) */
    int nc = orc_ccl(bin, H, W, 8, labels, stats, cap);               /* :26 */
    uint8_t* fill = (uint8_t*)calloc((size_t)nc + 1, 1);
    for (int i = 1; i <= nc; ++i) {
        int32_t* s = stats + (size_t)i * 5;
        if (s[0] > 0 && s[1] > 0 && s[2] < W - 1 && s[3] < H - 1 && s[4] < min_area) fill[i] = 1;  /* :40 */
    }
    for (size_t i = 0; i < n; ++i)
        if (labels[i] && fill[labels[i]]) mask[i] = (uint8_t)fg;      /* :41 */
    free(fill);

    for (size_t i = 0; i < n; ++i) bin[i] = mask[i] == fg ? 255 : 0;   /* :57 */
    morph3(bin, tmp, H, W, 1);                                         /* :60 OPEN = erode ... */
    morph3(tmp, bin, H, W, 0);                                         /*          ... then dilate */
    nc = orc_ccl(bin, H, W, 8, labels, stats, cap);                   /* :64 */
    for (size_t i = 0; i < n; ++i) {
        int l = labels[i];
        out[i] = (l && stats[(size_t)l * 5 + 4] >= min_area) ? (uint8_t)fg : 0;   /* :66-76 */
    }
    free(mask); free(bin); free(tmp); free(labels); free(stats);
}

/* ------------------------------------------------------------------ mask2polygon (src/mask2polygon.cpp) */

static int fgpix(const uint8_t* m, int H, int W, int thr, int x, int y) {
    if (x < 0 || y < 0 || x >= W || y >= H) return 0;
    return m[(size_t)y * W + x] > thr;                                 /* :31 threshold(127): v > 127 */
}

/* src/mask2polygon.cpp:29-36: threshold + findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE).
 * Output CSR: contour c = xy[2*cstart[c] .. 2*cstart[c+1]).  Returns the number of contours;
 * *n_pts receives the total point count.  If a capacity is too small nothing beyond it is written
 * but counting continues (call again with larger buffers). */
int orc_find_contours(const uint8_t* mask, int H, int W, int thr, int32_t* xy, int64_t cap_pts, int32_t* cstart,
                      int cap_c, int64_t* n_pts) {
    size_t n = (size_t)H * W;
    uint8_t* fg = (uint8_t*)malloc(n);
    for (size_t i = 0; i < n; ++i) fg[i] = mask[i] > thr;
    /* (1) 8-connected foreground components; label order = raster order of first pixel */
    int32_t* lab = (int32_t*)malloc(n * sizeof(int32_t));
    int ncomp = orc_ccl(fg, H, W, 8, lab, NULL, 0);
    int32_t* first = (int32_t*)malloc(((size_t)ncomp + 1) * sizeof(int32_t));
    for (int i = 0; i <= ncomp; ++i) first[i] = -1;
    for (size_t i = 0; i < n; ++i)
        if (lab[i] && first[lab[i]] < 0) first[lab[i]] = (int32_t)i;
    /* (2) 4-connected background on the image padded by a 1-px zero frame; frame component = label of (0,0) */
    int PW = W + 2, PH = H + 2;
    uint8_t* bg = (uint8_t*)malloc((size_t)PW * PH);
    for (int y = 0; y < PH; ++y)
        for (int x = 0; x < PW; ++x) bg[(size_t)y * PW + x] = !fgpix(mask, H, W, thr, x - 1, y - 1);
    int32_t* blab = (int32_t*)malloc((size_t)PW * PH * sizeof(int32_t));
    orc_ccl(bg, PH, PW, 4, blab, NULL, 0);
    const int frame = blab[0];

    int nout = 0;
    int64_t np = 0;
    /* (3) order: descending raster order of the start pixel */
    for (int c = ncomp; c >= 1; --c) {
        int s = first[c];
        int sx = s % W, sy = s / W;
        if (blab[(size_t)(sy + 1) * PW + (sx - 1 + 1)] != frame) continue;   /* left neighbour must be frame background */
        if (nout < cap_c) cstart[nout] = (int32_t)np;
        /* (5) find the LAST pixel: clockwise from W exclusive: NW,N,NE,E,SE,S,SW,W = 3,2,1,0,7,6,5,4 */
        int dL = -1;
        for (int k = 0; k < 8; ++k) {
            int d = (3 - k) & 7;
            if (fgpix(mask, H, W, thr, sx + DX8[d], sy + DY8[d])) { dL = d; break; }
        }
        if (dL < 0) {                                                   /* (6) single pixel */
            if (np < cap_pts) { xy[2 * np] = sx; xy[2 * np + 1] = sy; }
            ++np; ++nout;
            continue;
        }
        int lx = sx + DX8[dL], ly = sy + DY8[dL];
        int px = sx, py = sy, dprev = dL;
        int prev_out = (dL + 4) & 7;                                    /* d_out of L points at the start */
        for (;;) {
            int d = dprev;
            int qx, qy;
            do {                                                        /* counter-clockwise from dprev+1 */
                d = (d + 1) & 7;
                qx = px + DX8[d]; qy = py + DY8[d];
            } while (!fgpix(mask, H, W, thr, qx, qy));
            if (d != prev_out) {                                        /* (6) CHAIN_APPROX_SIMPLE */
                if (np < cap_pts) { xy[2 * np] = px; xy[2 * np + 1] = py; }
                ++np;
                prev_out = d;
            }
            if (qx == sx && qy == sy && px == lx && py == ly) break;
            px = qx; py = qy; dprev = (d + 4) & 7;
        }
        ++nout;
    }
    if (nout <= cap_c) cstart[nout] = (int32_t)np;   /* cstart holds cap_c + 1 entries */
    *n_pts = np;
    free(fg); free(lab); free(first); free(bg); free(blab);
    return nout;
}

/* src/mask2polygon.cpp:54-55 */
void orc_map_points(int32_t* xy, int64_t n_pts, double scale_x, double scale_y) {
    for (int64_t i = 0; i < n_pts; ++i) {
        xy[2 * i] = (int)(xy[2 * i] * scale_x);
        xy[2 * i + 1] = (int)(xy[2 * i + 1] * scale_y);
    }
}
