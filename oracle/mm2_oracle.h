/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the four hot-path stages of the AirLift minimap2 fork
 * (SURVEY.md §8a), written to be read next to the reference, each function
 * citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may link or call this.
 * The product (airlift_b200/) never does.
 *
 * Parity status: PINNED -- tests/test_oracle_vs_ref.py checks every function
 * here against the compiled reference (oracle/_ref/libmm2ref.so, built from
 * /root/reference by oracle/Makefile) and against committed fixtures in
 * tests/golden/ that were generated from that same reference build.
 */
#ifndef MM2_ORACLE_H
#define MM2_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t x, y; } orc128_t;

/* sketch.c:77-143 */
int orc_sketch(const char *str, int len, int w, int k, uint32_t rid, int is_hpc, orc128_t *out, int cap);

/* ksort.h:101-151 instantiated at misc.c:155-159 */
void orc_radix_sort_128x(orc128_t *beg, orc128_t *end);
void orc_radix_sort_64(uint64_t *beg, uint64_t *end);

/* index.c:191-243 (build) + index.c:81-98 (lookup): a sorted-table stand-in that
 * returns exactly what mm_idx_get returns: n and the positions in ascending order */
typedef struct orc_idx_s orc_idx_t;
orc_idx_t *orc_idx_build(int w, int k, int is_hpc, int n_seq, const char **seqs);
void orc_idx_destroy(orc_idx_t *mi);
const uint64_t *orc_idx_get(const orc_idx_t *mi, uint64_t minier, int *n);
int32_t orc_idx_cal_max_occ(const orc_idx_t *mi, float f); /* index.c:164-185 */
int64_t orc_idx_n_minimizers(const orc_idx_t *mi);

/* map.c:90-247: query minimizers -> sorted anchors.  heap_sort selects
 * collect_seed_hits_heap (map.c:149) vs collect_seed_hits (map.c:215).
 * flag carries only MM_F_FOR_ONLY/MM_F_REV_ONLY (skip_seed, map.c:139-145).
 * Returns n_a; a[] must hold the sum of occurrences (call with a==NULL to size). */
int64_t orc_collect_seeds(const orc_idx_t *mi, int heap_sort, int64_t flag, int max_occ, int n_mv, const orc128_t *mv,
                          int qlen, orc128_t *a, int *rep_len, int *n_mini_pos, uint64_t *mini_pos);

/* chain.c:22-162.  a[] (n entries) is overwritten with the compacted chains;
 * u[] receives score<<32|cnt per chain.  Returns n_u. */
int orc_chain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                 int is_cdna, int n_segs, int64_t n, orc128_t *a, uint64_t *u);

/* ksw2.h:23-32 */
typedef struct {
	uint32_t max; int zdropped;
	int max_q, max_t, mqe, mqe_t, mte, mte_q, score;
	int m_cigar, n_cigar, reach_end;
	uint32_t *cigar; /* malloc'd; caller frees */
} orc_extz_t;

/* ksw2_extd2_sse.c:19-393 (the SSE4.1 instantiation), lane-by-lane */
void orc_ksw_extd2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
                   int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus, int flag, orc_extz_t *ez);
/* DP cells inside the true band (the GCUPS unit) */
int64_t orc_ksw_band_cells(int qlen, int tlen, int w);

#ifdef __cplusplus
}
#endif
#endif
