/* dev analysis: how far does a speculative decode (guess: block of component c starts at bit i*S) run before it merges
 * with the true parse?  gcc -O2 -o tests/tools/jpeg_sync_sim tests/tools/jpeg_sync_sim.c && ./jpeg_sync_sim file.jpg [S] */
#include <stdio.h>
#include "../../oracle/jpeg_ref.c"

static uint8_t* U; static size_t NU;   /* unstuffed stream */
static uint32_t peekU(size_t p) { uint64_t v = 0; size_t b = p >> 3; for (int k = 0; k < 8; ++k) v = (v << 8) | (b + k < NU ? U[b + k] : 0xFF); return (uint32_t)((v << (p & 7)) >> 32); }
typedef struct { size_t p; int z, c; } St;
static const Jpeg* J; static int bpm, comp_of[8];
static int step(St* s) {   /* one symbol; returns 1 if a block completed */
    uint32_t w = peekU(s->p); int comp = comp_of[s->c];
    const Huff* t = s->z ? &J->ac[J->ta[comp]] : &J->dc[J->td[comp]];
    int len = 16, sym = 0;
    for (int l = 1; l <= 16; ++l) { int code = w >> (32 - l); if (t->maxcode[l] >= 0 && code <= t->maxcode[l] && code >= t->mincode[l]) { len = l; sym = t->vals[t->valptr[l] + code - t->mincode[l]]; break; } }
    int sz = sym & 15; s->p += len + sz;
    if (s->z == 0) s->z = 1; else { int r = sym >> 4; if (sz) s->z += r + 1; else s->z = (r == 15) ? s->z + 16 : 64; }
    if (s->z >= 64) { s->z = 0; if (++s->c == bpm) s->c = 0; return 1; }
    return 0;
}
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t* d = malloc(n); fread(d, 1, n, f); Jpeg j; if (parse(d, n, &j)) return 1; J = &j;
    size_t S = argc > 2 ? atoi(argv[2]) : 1024;
    U = malloc(j.scan_len); NU = 0;
    for (size_t i = 0; i + 1 < j.scan_len; ++i) { if (j.scan[i] == 0xFF && j.scan[i + 1] == 0xD9) break; U[NU++] = j.scan[i]; if (j.scan[i] == 0xFF && j.scan[i + 1] == 0) ++i; }
    bpm = 0; for (int c = 0; c < j.ncomp; ++c) for (int k = 0; k < j.hs[c] * j.vs[c]; ++k) comp_of[bpm++] = c;
    size_t bits = NU * 8; uint16_t* truth = calloc(bits + 64, 2);   /* (z<<4|c)+1 at every true symbol start */
    St s = {0, 0, 0}; while (s.p + 8 <= bits) { truth[s.p] = (uint16_t)(((s.z << 4) | s.c) + 1); step(&s); }
    size_t nsub = (bits + S - 1) / S; long hist[64] = {0}; size_t worst = 0, worst_i = 0; double sum = 0; long hist6[64] = {0}; size_t worst6 = 0; double sum6 = 0;
    for (size_t i = 1; i < nsub; ++i) {
        size_t best = (size_t)-1;
        for (int h = 0; h < bpm; ++h) {
            St t = {i * S, 0, h}; size_t lim = i * S + 4000 * S;
            while (t.p + 8 <= bits && t.p < lim && truth[t.p] != (uint16_t)(((t.z << 4) | t.c) + 1)) step(&t);
            size_t dist = (t.p - i * S) / S;
            if (h == 0) { sum += dist; if (dist > worst) { worst = dist; worst_i = i; } hist[dist < 63 ? dist : 63]++; }
            if (dist < best) best = dist;
        }
        sum6 += best; if (best > worst6) worst6 = best; hist6[best < 63 ? best : 63]++;
    }
    printf("S=%zu nsub=%zu  c=0 guess: mean %.2f worst %zu (at sub %zu)   best-of-%d: mean %.2f worst %zu\n", S, nsub, sum / nsub, worst, worst_i, bpm, sum6 / nsub, worst6);
    printf("hist c=0  :"); for (int k = 0; k < 64; ++k) if (hist[k]) printf(" %d:%ld", k, hist[k]); printf("\nhist best :"); for (int k = 0; k < 64; ++k) if (hist6[k]) printf(" %d:%ld", k, hist6[k]); printf("\n");
    return 0;
}
