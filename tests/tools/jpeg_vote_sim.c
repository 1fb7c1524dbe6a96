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
    size_t nsub = (bits + S - 1) / S;
    int span = argc > 3 ? atoi(argv[3]) : 1;     /* subsequences each hypothesis decodes in round 0 */
    St* tb = calloc(nsub + 2, sizeof(St)); { St t = {0,0,0}; for (size_t i = 0; i <= nsub; ++i) { while (t.p < i * S && t.p + 8 <= bits) step(&t); tb[i] = t; } }
    long ok_plural = 0, wrong_plural = 0, none = 0, h0_ok = 0, any_ok = 0; 
    char* good = calloc(nsub + 2, 1); good[0] = 1;
    for (size_t i = 0; i + span < nsub; ++i) {
        St e[8]; int cnt[8] = {0};
        for (int h = 0; h < bpm; ++h) { St t = {i * S, 0, h}; while (t.p < (i + span) * S && t.p + 8 <= bits) step(&t); e[h] = t; }
        int best = 0; for (int h = 0; h < bpm; ++h) { for (int g = 0; g < bpm; ++g) if (e[g].p == e[h].p && e[g].z == e[h].z && e[g].c == e[h].c) cnt[h]++; if (cnt[h] > cnt[best]) best = h; }
        St T = tb[i + span]; int is_ok = e[best].p == T.p && e[best].z == T.z && e[best].c == T.c;
        int anyok = 0; for (int h = 0; h < bpm; ++h) if (e[h].p == T.p && e[h].z == T.z && e[h].c == T.c) anyok = 1;
        any_ok += anyok; h0_ok += (e[0].p == T.p && e[0].z == T.z && e[0].c == T.c);
        if (cnt[best] >= 2) { if (is_ok) ok_plural++; else wrong_plural++; } else { none++; }
        good[i + span] = is_ok;
    }
    size_t run = 0, worst = 0, bad = 0; for (size_t i = 0; i < nsub; ++i) { if (!good[i] && i >= (size_t)span) { run++; bad++; if (run > worst) worst = run; } else run = 0; }
    printf("S=%zu span=%d nsub=%zu: plurality right %ld wrong %ld, no plurality %ld; h0 right %ld, any right %ld; bad boundaries %zu, longest bad run %zu\n", S, span, nsub, ok_plural, wrong_plural, none, h0_ok, any_ok, bad, worst);
    return 0;
}
