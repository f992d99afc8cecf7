/* TEST INFRASTRUCTURE ONLY — command-line front end of the CPU oracle.
 * Mirrors `orbit <x> <y> 0` (orbit.cpp:26-66) and prints the reference's three
 * timing lines (orbit.cpp:284-286: CountCopy/Partition in microseconds,
 * MakeAxis in milliseconds) followed by extra statistics on stderr.
 *
 * usage: orb_oracle <x> <y> [--ties=hoare|canonical] [--full] [--shards=N]
 *                   [--threads=N] [--trace=FILE] [--trace-particles] [--tight-box]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "orb_oracle.h"

int main(int argc, char **argv) {
    if (argc < 3) { printf("Usage: %s <N> <d>\n", argv[0]); return 1; }   /* orbit.cpp:26-29 */
    int pN = (int)strtol(argv[1], NULL, 0), pd = (int)strtol(argv[2], NULL, 0);
    uint64_t N = 1ULL << pN;
    orb_oracle_params p;
    memset(&p, 0, sizeof(p));
    p.d = 1 << pd; p.ties = ORB_TIES_HOARE; p.n_shards = 1; p.n_threads = 1; p.max_iter = 32;
    for (int i = 3; i < argc; ++i) {
        if (!strcmp(argv[i], "--ties=canonical")) p.ties = ORB_TIES_CANONICAL;
        else if (!strcmp(argv[i], "--ties=hoare")) p.ties = ORB_TIES_HOARE;
        else if (!strcmp(argv[i], "--full")) p.full_levels = 1;
        else if (!strcmp(argv[i], "--tight-box")) p.tight_box = 1;
        else if (!strncmp(argv[i], "--shards=", 9)) p.n_shards = atoi(argv[i] + 9);
        else if (!strncmp(argv[i], "--threads=", 10)) p.n_threads = atoi(argv[i] + 10);
        else if (!strncmp(argv[i], "--trace=", 8)) p.trace_path = argv[i] + 8;
        else if (!strcmp(argv[i], "--trace-particles")) p.trace_particles = 1;
    }
    /* orbit.cpp:83: every thread generates N/Threads particles; one stream, consecutive slices */
    uint64_t nLocal = N / (uint64_t)p.n_shards, nTot = nLocal * (uint64_t)p.n_shards;
    float *x = malloc(nTot * 4), *y = malloc(nTot * 4), *z = malloc(nTot * 4);
    uint64_t *off = malloc(((size_t)p.n_shards + 1) * 8);
    for (int s = 0; s <= p.n_shards; ++s) off[s] = nLocal * (uint64_t)s;
    orb_xorshf96_state g;
    orb_oracle_xorshf96_init(&g);
    orb_oracle_generate_uniform(&g, x, y, z, nTot);
    size_t nHeap = 2 * (size_t)p.d - 1;
    orb_oracle_cell *heap = malloc(nHeap * sizeof(orb_oracle_cell));
    uint32_t *ranges = malloc((size_t)p.n_shards * nHeap * 2 * 4);
    orb_oracle_stats st;
    int rc = orb_oracle_build(&p, x, y, z, off, heap, ranges, &st);
    if (rc) { fprintf(stderr, "orb_oracle_build failed: %d\n", rc); return 2; }
    printf("CountCopy-%u-%u, %u \n", pN, pd, (unsigned)(st.t_count_s * 1e6));
    printf("Partition-%u-%u, %u \n", pN, pd, (unsigned)(st.t_partition_s * 1e6));
    printf("MakeAxis-%u-%u, %u \n", pN, pd, (unsigned)(st.t_makeaxis_s * 1e3));
    fprintf(stderr, "levels %d total_s %.6f passes/N %.4f ties %llu iters", st.n_levels, st.t_total_s,
            (double)st.active_passes / (double)nTot, (unsigned long long)st.tie_particles);
    for (int l = 0; l < st.n_levels; ++l) fprintf(stderr, " %d", st.iters[l]);
    fprintf(stderr, "\n");
    /* rangeHash of SURVEY.md Appendix B: FNV-1a-style over the `end` index of each last-level cell in id order */
    int lastL = st.n_levels;
    unsigned long long h = 1469598103934665603ULL;
    for (int id = (1 << lastL) - 1; id <= (1 << (lastL + 1)) - 2; ++id) h = (h ^ ranges[2 * id + 1]) * 1099511628211ULL;
    fprintf(stderr, "rangeHash %016llx\n", h);
    /* heapHash: FNV-1a over the bytes of the whole cell heap (padding bytes are zero) */
    fprintf(stderr, "heapHash %016llx\n", (unsigned long long)orb_oracle_fnv1a(heap, nHeap * sizeof(orb_oracle_cell)));
    return 0;
}
