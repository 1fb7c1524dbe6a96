/* oracle/jpeg_ref.c -- TEST INFRASTRUCTURE ONLY (checker + cpu_baseline; never linked into the product).
 *
 * CPU restatement of what the reference does to a camera frame before the hot path:
 * `cv::imread(path)` (samples/main.cpp:24-40 of the reference) -> OpenCV's bundled libjpeg-turbo with its
 * defaults (JDCT_ISLOW, do_fancy_upsampling = TRUE, YCbCr -> BGR).  libjpeg-turbo is a third-party
 * dependency that is not vendored in /root/reference, so this file restates its published algorithm
 * (ITU-T T.81 baseline Huffman decoding; the 13-bit fixed-point Loeffler-Ligtenberg-Moschytz inverse DCT of
 * jidctint.c; the triangle-filter "fancy" chroma upsampling of jdsample.c; the 16-bit fixed-point colour
 * conversion of jdcolor.c).  Pinned in tests/test_oracle_jpeg.py against cv2.imdecode on the reference's own
 * frames (assets/images/0.jpg, 5.jpg -> tests/golden/frames) and on synthetic files (4:4:4, 4:2:2, 4:2:0,
 * grayscale, restart intervals, odd sizes): bit-exact.
 *
 * Scope: baseline / extended sequential (SOF0 / SOF1), 8 bit, Huffman, one interleaved scan, 1 or 3 components,
 * luma sampling 1x1, 2x1 or 2x2 with 1x1 chroma.  Everything else returns a negative error code.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { JE_OK = 0, JE_TRUNC = -1, JE_UNSUPPORTED = -2, JE_CORRUPT = -3, JE_NOMEM = -4 };

static const uint8_t kNatural[64 + 16] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
    6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
    39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

typedef struct {
    int defined;
    uint8_t bits[17];
    uint8_t vals[256];
    int mincode[17], maxcode[18], valptr[17];
} Huff;

typedef struct {
    int w, h, ncomp;
    int hs[3], vs[3], tq[3], td[3], ta[3];
    uint16_t q[4][64];   /* natural order */
    int qdef[4];
    Huff dc[4], ac[4];
    int restart;
    int comp_id[3];
    int adobe_transform;   /* -1: no APP14 Adobe marker */
    int orientation;       /* EXIF tag 0x0112, 0 if absent */
    int jfif;              /* APP0 JFIF marker seen: YCbCr by definition */
    const uint8_t* scan;
    size_t scan_len;
} Jpeg;

static int build_huff(Huff* t) {
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        t->valptr[l] = k;
        t->mincode[l] = code;
        code += t->bits[l];
        k += t->bits[l];
        t->maxcode[l] = t->bits[l] ? code - 1 : -1;
        if (code > (1 << l)) return JE_CORRUPT;
        code <<= 1;
    }
    t->maxcode[17] = 0x7fffffff;
    t->defined = 1;
    return JE_OK;
}

static int parse(const uint8_t* d, size_t n, Jpeg* j) {
    memset(j, 0, sizeof(*j));
    j->adobe_transform = -1;
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return JE_CORRUPT;
    size_t i = 2;
    int have_sof = 0;
    while (i + 4 <= n) {
        if (d[i] != 0xFF) return JE_CORRUPT;
        while (i < n && d[i] == 0xFF) ++i;   /* fill bytes */
        if (i >= n) return JE_TRUNC;
        const int m = d[i++];
        if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xD9) return JE_CORRUPT;
        if (i + 2 > n) return JE_TRUNC;
        const size_t len = ((size_t)d[i] << 8) | d[i + 1];
        if (len < 2 || i + len > n) return JE_TRUNC;
        const uint8_t* p = d + i + 2;
        const size_t pl = len - 2;
        if (m == 0xDB) {
            size_t o = 0;
            while (o < pl) {
                const int prec = p[o] >> 4, id = p[o] & 15;
                ++o;
                if (id > 3 || o + (prec ? 128 : 64) > pl) return JE_CORRUPT;
                for (int k = 0; k < 64; ++k) {
                    const int v = prec ? ((p[o] << 8) | p[o + 1]) : p[o];
                    o += prec ? 2 : 1;
                    j->q[id][kNatural[k]] = (uint16_t)v;
                }
                j->qdef[id] = 1;
            }
        } else if (m == 0xC4) {
            size_t o = 0;
            while (o < pl) {
                if (o + 17 > pl) return JE_CORRUPT;
                const int tc = p[o] >> 4, id = p[o] & 15;
                if (tc > 1 || id > 3) return JE_CORRUPT;
                Huff* t = tc ? &j->ac[id] : &j->dc[id];
                int cnt = 0;
                t->bits[0] = 0;
                for (int l = 1; l <= 16; ++l) cnt += (t->bits[l] = p[o + l]);
                o += 17;
                if (cnt > 256 || o + cnt > pl) return JE_CORRUPT;
                memcpy(t->vals, p + o, cnt);
                o += cnt;
                const int st = build_huff(t);
                if (st) return st;
            }
        } else if (m == 0xC0 || m == 0xC1) {
            if (pl < 6) return JE_CORRUPT;
            if (p[0] != 8) return JE_UNSUPPORTED;
            j->h = (p[1] << 8) | p[2];
            j->w = (p[3] << 8) | p[4];
            j->ncomp = p[5];
            if (j->ncomp != 1 && j->ncomp != 3) return JE_UNSUPPORTED;
            if (pl < (size_t)(6 + 3 * j->ncomp) || j->w == 0 || j->h == 0) return JE_CORRUPT;
            for (int c = 0; c < j->ncomp; ++c) {
                j->comp_id[c] = p[6 + 3 * c];
                j->hs[c] = p[7 + 3 * c] >> 4;
                j->vs[c] = p[7 + 3 * c] & 15;
                j->tq[c] = p[8 + 3 * c];
                if (j->tq[c] > 3) return JE_CORRUPT;
            }
            have_sof = 1;
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC8 && m != 0xCC) {
            return JE_UNSUPPORTED;   /* progressive, lossless, arithmetic */
        } else if (m == 0xE1) {   /* EXIF orientation: cv::imread rotates / mirrors the decoded frame for values 2..8 */
            if (pl >= 14 && memcmp(p, "Exif\0\0", 6) == 0) {
                const uint8_t* t = p + 6;
                const size_t tn = pl - 6;
                const int le = t[0] == 'I' && t[1] == 'I', be = t[0] == 'M' && t[1] == 'M';
#define U16(o) (le ? (t[o] | (t[(o) + 1] << 8)) : ((t[o] << 8) | t[(o) + 1]))
#define U32(o) (le ? ((size_t)t[o] | ((size_t)t[(o) + 1] << 8) | ((size_t)t[(o) + 2] << 16) | ((size_t)t[(o) + 3] << 24)) \
                   : (((size_t)t[o] << 24) | ((size_t)t[(o) + 1] << 16) | ((size_t)t[(o) + 2] << 8) | (size_t)t[(o) + 3]))
                if ((le || be) && U16(2) == 42) {
                    const size_t ifd = U32(4);
                    if (ifd + 2 <= tn) {
                        const int entries = U16(ifd);
                        for (int e = 0; e < entries && ifd + 2 + (size_t)(e + 1) * 12 <= tn; ++e) {
                            const size_t o = ifd + 2 + (size_t)e * 12;
                            if (U16(o) == 0x0112) j->orientation = U16(o + 8);
                        }
                    }
                }
#undef U16
#undef U32
            }
        } else if (m == 0xE0) {
            if (pl >= 5 && memcmp(p, "JFIF", 5) == 0) j->jfif = 1;
        } else if (m == 0xEE) {
            if (pl >= 12 && memcmp(p, "Adobe", 5) == 0) j->adobe_transform = p[11];
        } else if (m == 0xDD) {
            if (pl < 2) return JE_CORRUPT;
            j->restart = (p[0] << 8) | p[1];
        } else if (m == 0xDA) {
            if (!have_sof || pl < 1 || p[0] != j->ncomp || pl < (size_t)(4 + 2 * j->ncomp)) return JE_UNSUPPORTED;
            for (int c = 0; c < j->ncomp; ++c) {
                j->td[c] = p[2 + 2 * c] >> 4;
                j->ta[c] = p[2 + 2 * c] & 15;
                if (j->td[c] > 3 || j->ta[c] > 3) return JE_CORRUPT;
            }
            j->scan = d + i + len;
            j->scan_len = n - (i + len);
            break;
        }
        i += len;
    }
    if (!j->scan) return JE_TRUNC;
    if (j->ncomp == 1) {
        j->hs[0] = j->vs[0] = 1;   /* a single-component scan is never interleaved: blocks in raster order */
    } else {
        if (j->hs[1] != 1 || j->vs[1] != 1 || j->hs[2] != 1 || j->vs[2] != 1) return JE_UNSUPPORTED;
        if (!((j->hs[0] == 1 && j->vs[0] == 1) || (j->hs[0] == 2 && j->vs[0] == 1) || (j->hs[0] == 2 && j->vs[0] == 2)))
            return JE_UNSUPPORTED;
    }
    for (int c = 0; c < j->ncomp; ++c)
        if (!j->qdef[j->tq[c]] || !j->dc[j->td[c]].defined || !j->ac[j->ta[c]].defined) return JE_CORRUPT;
    if (j->orientation > 1 && j->orientation <= 8) return JE_UNSUPPORTED;        /* cv::imread would rotate */
    if (j->ncomp == 3 && !j->jfif && (j->adobe_transform == 0 ||                  /* RGB-coded, no YCbCr transform */
                          (j->adobe_transform < 0 && j->comp_id[0] == 'R' && j->comp_id[1] == 'G' && j->comp_id[2] == 'B')))
        return JE_UNSUPPORTED;
    return JE_OK;
}

/* ---- entropy decoding (T.81 F.2.2; jdhuff.c decode_mcu_slow) ---- */
typedef struct {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc;
    int nbits;
    int marker;
} Bits;

static void fill(Bits* b) {
    while (b->nbits <= 56) {
        int c = 0;
        if (!b->marker && b->p < b->end) {
            c = *b->p;
            if (c == 0xFF) {
                if (b->p + 1 < b->end && b->p[1] == 0x00) {
                    b->p += 2;
                } else {
                    b->marker = 1;   /* RSTn / EOI: feed zeros until the caller resynchronises */
                    c = 0;
                }
            } else {
                ++b->p;
            }
        }
        b->acc |= (uint64_t)c << (56 - b->nbits);
        b->nbits += 8;
    }
}
static inline int peek(Bits* b, int n) { return (int)(b->acc >> (64 - n)); }
static inline void skip(Bits* b, int n) { b->acc <<= n; b->nbits -= n; }
static inline int getbits(Bits* b, int n) {
    if (n == 0) return 0;
    if (b->nbits < n) fill(b);
    const int v = peek(b, n);
    skip(b, n);
    return v;
}
static int decode_sym(Bits* b, const Huff* t) {
    if (b->nbits < 16) fill(b);
    const int look = peek(b, 16);
    for (int l = 1; l <= 16; ++l) {
        const int code = look >> (16 - l);
        if (t->maxcode[l] >= 0 && code <= t->maxcode[l] && code >= t->mincode[l]) {
            skip(b, l);
            return t->vals[t->valptr[l] + code - t->mincode[l]];
        }
    }
    return -1;
}
static inline int extend(int r, int s) { return r < (1 << (s - 1)) ? r - (1 << s) + 1 : r; }

/* coefficient blocks in scan (decode) order, each block in natural order, quantised values */
static int decode_scan(const Jpeg* j, int16_t* coef, long nblocks_expected) {
    const int mcu_w = 8 * j->hs[0], mcu_h = 8 * j->vs[0];
    const int mcus_x = (j->w + mcu_w - 1) / mcu_w, mcus_y = (j->h + mcu_h - 1) / mcu_h;
    int blocks_in[3], bpm = 0;
    for (int c = 0; c < j->ncomp; ++c) bpm += (blocks_in[c] = j->hs[c] * j->vs[c]);
    if ((long)mcus_x * mcus_y * bpm != nblocks_expected) return JE_CORRUPT;
    Bits b = {j->scan, j->scan + j->scan_len, 0, 0, 0};
    int pred[3] = {0, 0, 0};
    int16_t* out = coef;
    long mcu = 0;
    const long total = (long)mcus_x * mcus_y;
    int next_rst = 0;
    for (; mcu < total; ++mcu) {
        if (j->restart && mcu && mcu % j->restart == 0) {
            /* byte-align, expect RSTn */
            b.acc = 0;
            b.nbits = 0;
            if (!b.marker) {
                /* the decoder has not run into the marker yet: it sits at the next 0xFF */
                while (b.p < b.end && !(b.p[0] == 0xFF && b.p + 1 < b.end && b.p[1] >= 0xD0 && b.p[1] <= 0xD7)) ++b.p;
            }
            if (b.p + 1 >= b.end || b.p[0] != 0xFF || b.p[1] != 0xD0 + next_rst) return JE_CORRUPT;
            b.p += 2;
            b.marker = 0;
            next_rst = (next_rst + 1) & 7;
            pred[0] = pred[1] = pred[2] = 0;
        }
        for (int c = 0; c < j->ncomp; ++c) {
            const Huff* dc = &j->dc[j->td[c]];
            const Huff* ac = &j->ac[j->ta[c]];
            for (int k = 0; k < blocks_in[c]; ++k, out += 64) {
                memset(out, 0, 64 * sizeof(int16_t));
                int s = decode_sym(&b, dc);
                if (s < 0 || s > 15) return JE_CORRUPT;
                if (s) pred[c] += extend(getbits(&b, s), s);
                out[0] = (int16_t)pred[c];
                for (int z = 1; z < 64; ++z) {
                    int rs = decode_sym(&b, ac);
                    if (rs < 0) return JE_CORRUPT;
                    const int r = rs >> 4;
                    s = rs & 15;
                    if (s) {
                        z += r;
                        const int v = extend(getbits(&b, s), s);
                        out[kNatural[z]] = (int16_t)v;
                    } else {
                        if (r != 15) break;
                        z += 15;
                    }
                }
            }
        }
    }
    return JE_OK;
}

/* ---- jidctint.c (JDCT_ISLOW) on one dequantised block -> 8x8 samples ---- */
#define CONST_BITS 13
#define PASS1_BITS 2
#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))
static inline uint8_t clamp_sample(int v) {
    /* range_limit[(v) & RANGE_MASK] with the table centred on CENTERJSAMPLE: 10-bit wrap, then clamp */
    v = (v + 128) & 1023;
    if (v >= 512 + 256) return 0;      /* wrapped negatives */
    if (v > 255) return 255;
    return (uint8_t)v;
}
static void idct_islow(const int16_t* coef, const uint16_t* q, uint8_t* out, int stride) {
    int ws[64];
    for (int c = 0; c < 8; ++c) {
        int in[8];
        for (int r = 0; r < 8; ++r) in[r] = coef[r * 8 + c] * q[r * 8 + c];
        int z2 = in[2], z3 = in[6];
        int z1 = (z2 + z3) * 4433;
        int tmp2 = z1 + z3 * (-15137), tmp3 = z1 + z2 * 6270;
        z2 = in[0];
        z3 = in[4];
        int tmp0 = (z2 + z3) * (1 << CONST_BITS), tmp1 = (z2 - z3) * (1 << CONST_BITS);
        const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = in[7];
        tmp1 = in[5];
        tmp2 = in[3];
        tmp3 = in[1];
        z1 = tmp0 + tmp3;
        z2 = tmp1 + tmp2;
        z3 = tmp0 + tmp2;
        int z4 = tmp1 + tmp3;
        const int z5 = (z3 + z4) * 9633;
        tmp0 *= 2446;
        tmp1 *= 16819;
        tmp2 *= 25172;
        tmp3 *= 12299;
        z1 *= -7373;
        z2 *= -20995;
        z3 *= -16069;
        z4 *= -3196;
        z3 += z5;
        z4 += z5;
        tmp0 += z1 + z3;
        tmp1 += z2 + z4;
        tmp2 += z2 + z3;
        tmp3 += z1 + z4;
        ws[0 * 8 + c] = DESCALE(tmp10 + tmp3, CONST_BITS - PASS1_BITS);
        ws[7 * 8 + c] = DESCALE(tmp10 - tmp3, CONST_BITS - PASS1_BITS);
        ws[1 * 8 + c] = DESCALE(tmp11 + tmp2, CONST_BITS - PASS1_BITS);
        ws[6 * 8 + c] = DESCALE(tmp11 - tmp2, CONST_BITS - PASS1_BITS);
        ws[2 * 8 + c] = DESCALE(tmp12 + tmp1, CONST_BITS - PASS1_BITS);
        ws[5 * 8 + c] = DESCALE(tmp12 - tmp1, CONST_BITS - PASS1_BITS);
        ws[3 * 8 + c] = DESCALE(tmp13 + tmp0, CONST_BITS - PASS1_BITS);
        ws[4 * 8 + c] = DESCALE(tmp13 - tmp0, CONST_BITS - PASS1_BITS);
    }
    for (int r = 0; r < 8; ++r) {
        const int* w = ws + r * 8;
        int z2 = w[2], z3 = w[6];
        int z1 = (z2 + z3) * 4433;
        int tmp2 = z1 + z3 * (-15137), tmp3 = z1 + z2 * 6270;
        int tmp0 = (w[0] + w[4]) * (1 << CONST_BITS), tmp1 = (w[0] - w[4]) * (1 << CONST_BITS);
        const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7];
        tmp1 = w[5];
        tmp2 = w[3];
        tmp3 = w[1];
        z1 = tmp0 + tmp3;
        z2 = tmp1 + tmp2;
        z3 = tmp0 + tmp2;
        int z4 = tmp1 + tmp3;
        const int z5 = (z3 + z4) * 9633;
        tmp0 *= 2446;
        tmp1 *= 16819;
        tmp2 *= 25172;
        tmp3 *= 12299;
        z1 *= -7373;
        z2 *= -20995;
        z3 *= -16069;
        z4 *= -3196;
        z3 += z5;
        z4 += z5;
        tmp0 += z1 + z3;
        tmp1 += z2 + z4;
        tmp2 += z2 + z3;
        tmp3 += z1 + z4;
        uint8_t* o = out + r * stride;
        const int sh = CONST_BITS + PASS1_BITS + 3;
        o[0] = clamp_sample(DESCALE(tmp10 + tmp3, sh));
        o[7] = clamp_sample(DESCALE(tmp10 - tmp3, sh));
        o[1] = clamp_sample(DESCALE(tmp11 + tmp2, sh));
        o[6] = clamp_sample(DESCALE(tmp11 - tmp2, sh));
        o[2] = clamp_sample(DESCALE(tmp12 + tmp1, sh));
        o[5] = clamp_sample(DESCALE(tmp12 - tmp1, sh));
        o[3] = clamp_sample(DESCALE(tmp13 + tmp0, sh));
        o[4] = clamp_sample(DESCALE(tmp13 - tmp0, sh));
    }
}

/* ---- jdsample.c fancy upsampling + jdcolor.c ycc_rgb_convert ---- */
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

int jpeg_ref_info(const uint8_t* data, size_t n, int* w, int* h, int* ncomp, int* hs, int* vs, int* restart) {
    Jpeg j;
    const int st = parse(data, n, &j);
    if (st) return st;
    *w = j.w; *h = j.h; *ncomp = j.ncomp; *hs = j.hs[0]; *vs = j.vs[0]; *restart = j.restart;
    return JE_OK;
}

long jpeg_ref_num_blocks(const uint8_t* data, size_t n) {
    Jpeg j;
    const int st = parse(data, n, &j);
    if (st) return st;
    const int mcu_w = 8 * j.hs[0], mcu_h = 8 * j.vs[0];
    long bpm = 0;
    for (int c = 0; c < j.ncomp; ++c) bpm += j.hs[c] * j.vs[c];
    return (long)((j.w + mcu_w - 1) / mcu_w) * ((j.h + mcu_h - 1) / mcu_h) * bpm;
}

/* bgr: h*w*3 bytes (cv::imread layout); coef_out: optional, num_blocks*64 int16 in scan order */
int jpeg_ref_decode(const uint8_t* data, size_t n, uint8_t* bgr, int16_t* coef_out) {
    Jpeg j;
    int st = parse(data, n, &j);
    if (st) return st;
    const int H0 = j.hs[0], V0 = j.vs[0];
    const int mcu_w = 8 * H0, mcu_h = 8 * V0;
    const int mcus_x = (j.w + mcu_w - 1) / mcu_w, mcus_y = (j.h + mcu_h - 1) / mcu_h;
    int bpm = 0;
    for (int c = 0; c < j.ncomp; ++c) bpm += j.hs[c] * j.vs[c];
    const long nblocks = (long)mcus_x * mcus_y * bpm;
    int16_t* coef = coef_out ? coef_out : (int16_t*)malloc((size_t)nblocks * 64 * sizeof(int16_t));
    if (!coef) return JE_NOMEM;
    st = decode_scan(&j, coef, nblocks);
    uint8_t* plane[3] = {0, 0, 0};
    int pw[3], ph[3];
    if (!st) {
        for (int c = 0; c < j.ncomp; ++c) {
            pw[c] = mcus_x * 8 * j.hs[c];
            ph[c] = mcus_y * 8 * j.vs[c];
            plane[c] = (uint8_t*)malloc((size_t)pw[c] * ph[c]);
            if (!plane[c]) st = JE_NOMEM;
        }
    }
    if (!st) {
        const int16_t* b = coef;
        for (int my = 0; my < mcus_y; ++my)
            for (int mx = 0; mx < mcus_x; ++mx)
                for (int c = 0; c < j.ncomp; ++c)
                    for (int by = 0; by < j.vs[c]; ++by)
                        for (int bx = 0; bx < j.hs[c]; ++bx, b += 64) {
                            const int x = (mx * j.hs[c] + bx) * 8, y = (my * j.vs[c] + by) * 8;
                            idct_islow(b, j.q[j.tq[c]], plane[c] + (size_t)y * pw[c] + x, pw[c]);
                        }
        if (j.ncomp == 1) {
            for (int y = 0; y < j.h; ++y)
                for (int x = 0; x < j.w; ++x) {
                    const uint8_t v = plane[0][(size_t)y * pw[0] + x];
                    uint8_t* o = bgr + ((size_t)y * j.w + x) * 3;
                    o[0] = o[1] = o[2] = v;
                }
        } else {
            const int cw = (j.w * 1 + H0 - 1) / H0, ch = (j.h * 1 + V0 - 1) / V0;   /* downsampled_width / _height */
            const int fancy = cw > 2;   /* jinit_upsampler: fancy only when downsampled_width > 2 */
            for (int y = 0; y < j.h; ++y) {
                for (int x = 0; x < j.w; ++x) {
                    int cc[2];
                    for (int c = 1; c <= 2; ++c) {
                        const uint8_t* p = plane[c];
                        const int W = pw[c];
                        int v;
                        if (H0 == 1) {
                            v = p[(size_t)y * W + x];
                        } else if (V0 == 1) {          /* h2v1 */
                            const int i = x >> 1;
                            const uint8_t* row = p + (size_t)y * W;
                            if (!fancy) v = row[i];
                            else if (x & 1) v = (3 * row[i] + row[clampi(i + 1, 0, cw - 1)] + 2) >> 2;
                            else v = (3 * row[i] + row[clampi(i - 1, 0, cw - 1)] + 1) >> 2;
                        } else {                        /* h2v2 */
                            const int i = x >> 1, r = y >> 1;
                            if (!fancy) {
                                v = p[(size_t)r * W + i];
                            } else {
                                const int r1 = clampi((y & 1) ? r + 1 : r - 1, 0, ch - 1);
                                const uint8_t* row0 = p + (size_t)r * W;
                                const uint8_t* row1 = p + (size_t)r1 * W;
                                const int thisc = 3 * row0[i] + row1[i];
                                if (x & 1) {
                                    const int k = clampi(i + 1, 0, cw - 1);
                                    v = (3 * thisc + 3 * row0[k] + row1[k] + 7) >> 4;
                                } else {
                                    const int k = clampi(i - 1, 0, cw - 1);
                                    v = (3 * thisc + 3 * row0[k] + row1[k] + 8) >> 4;
                                }
                            }
                        }
                        cc[c - 1] = v;
                    }
                    const int Y = plane[0][(size_t)y * pw[0] + x], cb = cc[0] - 128, cr = cc[1] - 128;
                    const int R = Y + ((91881 * cr + 32768) >> 16);
                    const int G = Y + ((-22554 * cb - 46802 * cr + 32768) >> 16);
                    const int B = Y + ((116130 * cb + 32768) >> 16);
                    uint8_t* o = bgr + ((size_t)y * j.w + x) * 3;
                    o[0] = (uint8_t)clampi(B, 0, 255);
                    o[1] = (uint8_t)clampi(G, 0, 255);
                    o[2] = (uint8_t)clampi(R, 0, 255);
                }
            }
        }
    }
    for (int c = 0; c < 3; ++c) free(plane[c]);
    if (!coef_out) free(coef);
    return st;
}
