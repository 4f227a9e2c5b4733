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

/* bench.py's reference arm: the reference's own worker_pipeline (map.c:556-650), its three steps run back to back on one
 * thread instead of overlapped by kt_pipeline, so that step 1 -- kt_for(worker_for) -> mm_map_frag, the hot path -- can be
 * timed alone on all host threads.  Step 2 still formats and prints every record (the caller points stdout at /dev/null).
 * Returns the number of mini-batches done; t[3*i..3*i+2] = seconds of steps 0,1,2 of batch i; n[i] = its reads. */
int ref_pipeline_steps(const mm_idx_t *mi, const mm_mapopt_t *opt, int n_threads, int n_fp, const char **fn, int mini_batch_size,
                       int max_batches, double *t, long *n)
{
	pipeline_t pl;
	int done = 0;
	memset(&pl, 0, sizeof(pipeline_t));
	pl.n_fp = n_fp;
	pl.fp = open_bseqs(pl.n_fp, fn);
	if (pl.fp == 0) return -1;
	pl.opt = opt, pl.mi = mi;
	pl.n_threads = n_threads > 1? n_threads : 1;
	pl.mini_batch_size = mini_batch_size;
	while (done < max_batches) {
		double t0 = realtime(), t1, t2;
		step_t *s = (step_t*)worker_pipeline(&pl, 0, 0);
		if (s == 0) break;
		n[done] = s->n_seq;
		t1 = realtime();
		worker_pipeline(&pl, 1, s);
		t2 = realtime();
		worker_pipeline(&pl, 2, s);
		t[3 * done] = t1 - t0, t[3 * done + 1] = t2 - t1, t[3 * done + 2] = realtime() - t2;
		++done;
	}
	free(pl.str.s);
	{ int i; for (i = 0; i < pl.n_fp; ++i) mm_bseq_close(pl.fp[i]); }
	free(pl.fp);
	return done;
}
