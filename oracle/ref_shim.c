/* TEST HARNESS ONLY.  Compiles the reference's own map.c (in its Oracle-B form,
 * generated into _ref/patched/ by the Makefile) as part of this translation unit so
 * that its file-static stage functions can be called from the parity tests.  No
 * reference code is copied here; this file only adds exported trampolines. */
#include "patched/map.c"

mm128_t *ref_collect_seeds(const mm_idx_t *mi, int heap_sort, int64_t flag, int max_occ, int n_mv, const mm128_t *mv_a,
                           int qlen, int64_t *n_a, int *rep_len, int *n_mini_pos, uint64_t **mini_pos)
{
	mm_mapopt_t opt;
	mm128_v mv;
	memset(&opt, 0, sizeof(opt));
	opt.flag = flag;
	mv.n = mv.m = n_mv, mv.a = (mm128_t*)mv_a;
	if (heap_sort) return collect_seed_hits_heap(0, &opt, max_occ, mi, 0, &mv, qlen, n_a, rep_len, n_mini_pos, mini_pos);
	return collect_seed_hits(0, &opt, max_occ, mi, 0, &mv, qlen, n_a, rep_len, n_mini_pos, mini_pos);
}

/* mm_chain_dp frees its input with kfree(km,a); hand it a private copy */
int ref_chain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc, int is_cdna,
                 int n_segs, int64_t n, mm128_t *a_inout, uint64_t *u_out)
{
	mm128_t *a = (mm128_t*)malloc((n + 1) * sizeof(mm128_t)), *b;
	uint64_t *u = 0;
	int n_u = 0, i, n_v = 0;
	memcpy(a, a_inout, n * sizeof(mm128_t));
	b = mm_chain_dp(max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, is_cdna, n_segs, n, a, &n_u, &u, 0);
	for (i = 0; i < n_u; ++i) u_out[i] = u[i], n_v += (int32_t)u[i];
	if (b) memcpy(a_inout, b, n_v * sizeof(mm128_t));
	free(b); free(u);
	return n_u;
}

int ref_sizeof_mapopt(void) { return (int)sizeof(mm_mapopt_t); }
int ref_sizeof_reg1(void) { return (int)sizeof(mm_reg1_t); }
