/* Developer tool (CPU): the selection logic of the batched pruned FPS (csrc/fps_pruned.cu, "rounds") replayed on the host and
 * compared with the oracle's literal emulation of the reference kernel (oracle/pdab_oracle.c, orc_fps).  It answers two questions
 * before any GPU time is spent: is the acceptance rule exact (idx and temp identical), and how many samples does a round accept
 * under a given threshold controller.
 *
 *   gcc -O2 -ffp-contract=off -o /tmp/fps_batch_sim tools/fps_batch_sim.c oracle/pdab_oracle.c -lm
 *   /tmp/fps_batch_sim N m P NB T cloud seed       (cloud: 0 slab, 1 clustered, 2 all-equal, 3 grid with many ties)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void orc_fps(int b, int n, int m, const float *xyz, float *temp, int *idx);

static inline float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}
static unsigned brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static unsigned tie_key(int k, int L) {
    if (L == 0) return (unsigned)k;
    return brev((unsigned)k & ((1u << L) - 1u)) | ((unsigned)k >> L);
}
static unsigned spread10(unsigned v) {
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
static unsigned fbits(float f) {
    unsigned u;
    memcpy(&u, &f, 4);
    return u;
}
static int cmp_u64(const void *a, const void *b) {
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

typedef struct {
    uint64_t key;
    float x, y, z, sec;
} Cand;

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 16384, m = argc > 2 ? atoi(argv[2]) : 4096;
    const int P = argc > 3 ? atoi(argv[3]) : 16, NB = argc > 4 ? atoi(argv[4]) : 2, T = argc > 5 ? atoi(argv[5]) : 512;
    const int cloud = argc > 6 ? atoi(argv[6]) : 0;
    unsigned s = argc > 7 ? (unsigned)atoi(argv[7]) : 12345u;
    const int LO = argc > 8 ? atoi(argv[8]) : 6, HI = argc > 9 ? atoi(argv[9]) : 20, MAXC = argc > 10 ? atoi(argv[10]) : 32;
    const int CAP = P * NB * T;
    if (n > CAP) return printf("n > capacity\n"), 1;
    float *xyz = malloc(sizeof(float) * 3 * n);
#define RND() (s = s * 1664525u + 1013904223u, (s >> 8) * (1.0f / 16777216.0f))
    for (int i = 0; i < n; i++) {
        float x = RND() * 70.f, y = RND() * 80.f - 40.f, z = RND() * 4.f - 3.f;
        if (cloud == 1) {  /* a few dense clusters + background */
            if (i % 4) {
                const int c = i % 7;
                x = 10.f * c + RND() * 2.f, y = -30.f + 9.f * c + RND() * 2.f, z = RND();
            }
        } else if (cloud == 2) {
            x = 1.f, y = 2.f, z = 3.f;
        } else if (cloud == 3) {  /* lattice: masses of exactly equal distances */
            x = (float)(i % 32), y = (float)((i / 32) % 32), z = (float)(i / 1024);
        }
        xyz[3 * i] = x, xyz[3 * i + 1] = y, xyz[3 * i + 2] = z;
    }
    int L = 0;
    while ((1 << (L + 1)) <= n && L < 10) L++;

    /* oracle */
    float *tref = malloc(sizeof(float) * n), *d = malloc(sizeof(float) * n);
    int *iref = malloc(sizeof(int) * m), *idx = malloc(sizeof(int) * m);
    for (int i = 0; i < n; i++) tref[i] = d[i] = 1e10f;
    orc_fps(1, n, m, xyz, tref, iref);

    /* Morton ranks -> owner thread of every point */
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) {
            lo[a] = fminf(lo[a], xyz[3 * i + a]);
            hi[a] = fmaxf(hi[a], xyz[3 * i + a]);
        }
    const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    const float qs = ext > 0.f ? 1023.0f / ext : 0.f;
    uint64_t *keys = malloc(sizeof(uint64_t) * n);
    for (int i = 0; i < n; i++) {
        unsigned q[3];
        for (int a = 0; a < 3; a++) {
            q[a] = (unsigned)((xyz[3 * i + a] - lo[a]) * qs);
            if (q[a] > 1023u) q[a] = 1023u;
        }
        keys[i] = ((uint64_t)(spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2)) << 32) | (unsigned)i;
    }
    qsort(keys, n, sizeof(uint64_t), cmp_u64);
    int *owner = malloc(sizeof(int) * n);
    for (int r = 0; r < n; r++) owner[(unsigned)keys[r]] = (r / P) % T;

    /* batched rounds */
    uint64_t *tbest = malloc(sizeof(uint64_t) * T);
    int *tpos = malloc(sizeof(int) * T);
    float *tsec = malloc(sizeof(float) * T);
    Cand list[64];
    float px[64], py[64], pz[64];
    int npend = 1, it = 1;
    px[0] = xyz[0], py[0] = xyz[1], pz[0] = xyz[2];
    idx[0] = 0;
    float vref = 0.f, gap = 0.25f;
    long rounds = 0, fb_empty = 0, fb_over = 0, sum_nc = 0, hist[40] = {0}, blk_conf = 0, blk_sec = 0, blk_none = 0;
    while (it < m) {
        for (int r = 0; r < npend; r++)
            for (int k = 0; k < n; k++) d[k] = fminf(sqdist3(xyz[3 * k], xyz[3 * k + 1], xyz[3 * k + 2], px[r], py[r], pz[r]), d[k]);
        for (int t = 0; t < T; t++) tbest[t] = 0, tsec[t] = -1.f, tpos[t] = -1;
        for (int k = 0; k < n; k++) {
            const int t = owner[k];
            const uint64_t key = ((uint64_t)fbits(d[k]) << 32) | (unsigned)~tie_key(k, L);
            if (tpos[t] < 0 || key > tbest[t]) {
                if (tpos[t] >= 0) tsec[t] = fmaxf(tsec[t], d[tpos[t]]);
                tbest[t] = key, tpos[t] = k;
            } else {
                tsec[t] = fmaxf(tsec[t], d[k]);
            }
        }
        const float tau = vref > 0.f ? vref * (1.f - gap) : INFINITY;
        int nc = 0;
        for (int t = 0; t < T; t++)
            if (tpos[t] >= 0 && d[tpos[t]] >= tau) {
                if (nc < MAXC) {
                    const int k = tpos[t];
                    list[nc].key = tbest[t], list[nc].x = xyz[3 * k], list[nc].y = xyz[3 * k + 1], list[nc].z = xyz[3 * k + 2];
                    list[nc].sec = tsec[t];
                }
                nc++;
            }
        rounds++;
        int A;
        if (nc == 0 || nc > MAXC) {
            if (nc == 0) fb_empty++, gap = fminf(gap * 4.f, 0.5f);
            else fb_over++, gap = fmaxf(gap * 0.25f, 1e-7f);
            uint64_t best = 0;
            int bk = 0;
            for (int t = 0; t < T; t++)
                if (tpos[t] >= 0 && tbest[t] > best) best = tbest[t], bk = tpos[t];
            A = 1;
            idx[it] = bk;
            px[0] = xyz[3 * bk], py[0] = xyz[3 * bk + 1], pz[0] = xyz[3 * bk + 2];
            vref = d[bk];
        } else {
            sum_nc += nc;
            int rank[64], blocked[64];
            for (int i = 0; i < nc; i++) {
                float vi;
                const unsigned hb = (unsigned)(list[i].key >> 32);
                memcpy(&vi, &hb, 4);
                rank[i] = 0, blocked[i] = 0;
                for (int j = 0; j < nc; j++) {
                    if (list[j].key > list[i].key) {
                        rank[i]++;
                        const float dd = sqdist3(list[i].x, list[i].y, list[i].z, list[j].x, list[j].y, list[j].z);
                        if (dd < vi || list[j].sec >= vi) blocked[i] = 1;
                    }
                }
            }
            A = nc;
            for (int i = 0; i < nc; i++)
                if (blocked[i] && rank[i] < A) A = rank[i];
            if (A == nc) blk_none++;
            else
                for (int i = 0; i < nc; i++)
                    if (rank[i] == A) {   /* what stopped the round: a conflict or a hidden point? */
                        float vi;
                        const unsigned hb = (unsigned)(list[i].key >> 32);
                        memcpy(&vi, &hb, 4);
                        int conf = 0;
                        for (int j = 0; j < nc; j++)
                            if (list[j].key > list[i].key &&
                                sqdist3(list[i].x, list[i].y, list[i].z, list[j].x, list[j].y, list[j].z) < vi) conf = 1;
                        if (conf) blk_conf++; else blk_sec++;
                    }
            if (A > m - it) A = m - it;
            for (int i = 0; i < nc; i++)
                if (rank[i] < A) {
                    /* decode the original index from the tie key */
                    const unsigned tk = ~(unsigned)list[i].key;
                    int k = -1;
                    if (L == 0) k = (int)tk;
                    else {
                        const unsigned lowmask = (1u << (32 - L)) - 1u;
                        k = (int)(((tk & lowmask) << L) | brev(tk & ~lowmask));
                    }
                    idx[it + rank[i]] = k;
                    px[rank[i]] = list[i].x, py[rank[i]] = list[i].y, pz[rank[i]] = list[i].z;
                    if (rank[i] == A - 1) vref = d[k];
                }
            if (nc < LO) gap = fminf(gap * 1.5f, 0.5f);
            else if (nc > HI) gap = fmaxf(gap / 1.5f, 1e-7f);
        }
        hist[A]++;
        it += A;
        npend = it == m ? A - 1 : A;
    }
    for (int r = 0; r < npend; r++)
        for (int k = 0; k < n; k++) d[k] = fminf(sqdist3(xyz[3 * k], xyz[3 * k + 1], xyz[3 * k + 2], px[r], py[r], pz[r]), d[k]);
    int bad_i = 0, bad_t = 0;
    for (int i = 0; i < m; i++) bad_i += idx[i] != iref[i];
    for (int k = 0; k < n; k++) bad_t += memcmp(&d[k], &tref[k], 4) != 0;
    printf("n %d m %d P %d NB %d T %d cloud %d: idx mismatches %d, temp mismatches %d | rounds %ld (%.2f samples/round) fallbacks: empty %ld overflow %ld, mean list %.1f\n",
           n, m, P, NB, T, cloud, bad_i, bad_t, rounds, (double)(m - 1) / rounds, fb_empty, fb_over,
           rounds - fb_empty - fb_over ? (double)sum_nc / (rounds - fb_empty - fb_over) : 0.0);
    printf("  rounds ended by: list exhausted %ld, conflict %ld, hidden-point bound %ld\n", blk_none, blk_conf, blk_sec);
    printf("  accepted histogram:");
    for (int a = 1; a <= 32; a++)
        if (hist[a]) printf(" %d:%ld", a, hist[a]);
    printf("\n");
    return bad_i || bad_t;
}
