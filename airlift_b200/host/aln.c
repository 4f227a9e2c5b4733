/* aln.c -- base-level alignment of chained regions: the host half of K4.
 *
 * The reference walks each region with blocking DP calls (mm_align_skeleton -> mm_align1 ->
 * mm_align_pair -> ksw_extd2_sse, align.c:313-913).  Here the DP runs on the GPU in batches, so the
 * same walk is made RESUMABLE: every DP request goes through a per-segment job cache.  A request that
 * is not cached yet is queued and answered with an empty result; the walk continues (so that all
 * independent requests of the region are discovered in one pass) and the region is re-walked from its
 * saved state once the batch has run.  Window selection, CIGAR stitching, z-drop splitting and the
 * CIGAR post-processing below follow the reference line by line in behaviour (cited per function).
 */
#include <math.h>
#include <stdio.h>
#include "mm2b_priv.h"

typedef struct {  /* what mm_align1 reads from a ksw_extz_t (ksw2.h:23-32) */
	uint32_t max; int zdropped;
	int max_q, max_t, mqe, mqe_t, score, n_cigar, reach_end;
	const uint32_t *cigar;
} ez_t;

typedef struct {
	mm_alnseg_t *seg;
	const mm_mapopt_t *opt;
	const mm_idx_t *mi;
	int pending;       /* a DP result needed by the current region is not available yet */
	int8_t mat[25];
} walk_t;

MM_FN static void ez_reset(ez_t *ez)
{
	ez->max_q = ez->max_t = ez->mqe_t = -1;
	ez->max = 0, ez->score = ez->mqe = KSW_NEG_INF;
	ez->n_cigar = 0, ez->zdropped = 0, ez->reach_end = 0, ez->cigar = 0;
}

MM_FN static void gen_simple_mat(int8_t *mat, int a, int b, int sc_ambi)
{ /* align.c:9-22, m = 5 */
	int i, j;
	a = a < 0 ? -a : a, b = b > 0 ? -b : b, sc_ambi = sc_ambi > 0 ? -sc_ambi : sc_ambi;
	for (i = 0; i < 4; ++i) {
		for (j = 0; j < 4; ++j) mat[i * 5 + j] = (int8_t)(i == j ? a : b);
		mat[i * 5 + 4] = (int8_t)sc_ambi;
	}
	for (j = 0; j < 5; ++j) mat[20 + j] = (int8_t)sc_ambi;
}

/* ---- DP through the cache (replaces mm_align_pair, align.c:313-339) */

MM_FN static int dp_request(walk_t *w, int q_rev, int q_start, int q_len, int rid, int t_start, int t_len, int reversed,
                      int bw, int end_bonus, int zdrop, int flag, ez_t *ez)
{
	mm_dpcache_t *c = &w->seg->cache;
	const mm_mapopt_t *opt = w->opt;
	mm_dpjob_t *j;
	int i;
	ez_reset(ez);
	if (opt->max_sw_mat > 0 && (int64_t)t_len * q_len > opt->max_sw_mat) { ez->zdropped = 1; return 1; } /* align.c:323-325 */
	if (q_len <= 0 || t_len <= 0) return 1; /* ksw returns the reset state (ksw2_extd2_sse.c:68) */
	for (i = 0; i < c->n; ++i) {
		const mmg_ksw_job_t *k = &c->a[i].job;
		if (k->q_rev == q_rev && k->q_start == q_start && k->q_len == q_len && k->rid == rid && k->t_start == t_start && k->t_len == t_len
			&& k->reversed == reversed && k->w == bw && k->zdrop == zdrop && k->end_bonus == end_bonus && k->flag == flag) break;
	}
	if (i < c->n) {
		j = &c->a[i];
		if (!j->done) { w->pending = 1; return 0; }
		ez->max = j->ez.max, ez->zdropped = j->ez.zdropped, ez->max_q = j->ez.max_q, ez->max_t = j->ez.max_t;
		ez->mqe = j->ez.mqe, ez->mqe_t = j->ez.mqe_t, ez->score = j->ez.score, ez->n_cigar = j->ez.n_cigar, ez->reach_end = j->ez.reach_end;
		ez->cigar = j->cigar;
		return 1;
	}
	if (c->n == c->m) { const size_t old = (size_t)c->m * sizeof(mm_dpjob_t); c->m = c->m ? c->m << 1 : 8; c->a = (mm_dpjob_t*)mm_arealloc(c->a, old, (size_t)c->m * sizeof(mm_dpjob_t)); }
	j = &c->a[c->n++];
	memset(j, 0, sizeof(*j));
	j->job.seq_id = w->seg->seq_id, j->job.q_rev = q_rev, j->job.q_start = q_start, j->job.q_len = q_len;
	j->job.rid = rid, j->job.t_start = t_start, j->job.t_len = t_len, j->job.reversed = reversed;
	j->job.w = bw, j->job.zdrop = zdrop, j->job.end_bonus = end_bonus, j->job.flag = flag;
	w->pending = 1;
	return 0;
}

/* ---- CIGAR bookkeeping */

MM_FN static void append_cigar(mm_reg1_t *r, uint32_t n_cigar, const uint32_t *cigar)
{ /* align.c:288-311 */
	mm_extra_t *p;
	if (n_cigar == 0) return;
	if (r->p == 0) {
		uint32_t capacity = n_cigar + sizeof(mm_extra_t) / 4;
		capacity = mm_roundup32(capacity);
		r->p = (mm_extra_t*)calloc(capacity, 4);
		r->p->capacity = capacity;
	} else if (r->p->n_cigar + n_cigar + sizeof(mm_extra_t) / 4 > r->p->capacity) {
		r->p->capacity = r->p->n_cigar + n_cigar + sizeof(mm_extra_t) / 4;
		r->p->capacity = mm_roundup32(r->p->capacity);
		r->p = (mm_extra_t*)realloc(r->p, (size_t)r->p->capacity * 4);
	}
	p = r->p;
	if (p->n_cigar > 0 && (p->cigar[p->n_cigar - 1] & 0xf) == (cigar[0] & 0xf)) { /* merge at the seam */
		p->cigar[p->n_cigar - 1] += cigar[0] >> 4 << 4;
		if (n_cigar > 1) memcpy(p->cigar + p->n_cigar, cigar + 1, (size_t)(n_cigar - 1) * 4);
		p->n_cigar += n_cigar - 1;
	} else {
		memcpy(p->cigar + p->n_cigar, cigar, (size_t)n_cigar * 4);
		p->n_cigar += n_cigar;
	}
}

MM_FN static void fix_cigar(mm_reg1_t *r, const uint8_t *qseq, const uint8_t *tseq, int *qshift, int *tshift)
{ /* align.c:91-167: left-align indels, collapse I/D runs, drop zero-length ops and a leading I/D */
	mm_extra_t *p = r->p;
	int32_t toff = 0, qoff = 0, to_shrink = 0;
	uint32_t k;
	*qshift = *tshift = 0;
	if (p->n_cigar <= 1) return;
	for (k = 0; k < p->n_cigar; ++k) {
		const uint32_t op = p->cigar[k] & 0xf, len = p->cigar[k] >> 4;
		if (len == 0) to_shrink = 1;
		if (op == 0) toff += len, qoff += len;
		else if (op == 1 || op == 2) {
			if (k > 0 && k < p->n_cigar - 1 && (p->cigar[k-1] & 0xf) == 0 && (p->cigar[k+1] & 0xf) == 0) {
				const int prev_len = p->cigar[k-1] >> 4;
				const uint8_t *sq = op == 1 ? qseq : tseq;
				const int32_t o = op == 1 ? qoff : toff;
				int l;
				for (l = 0; l < prev_len; ++l)
					if (sq[o - 1 - l] != sq[o + len - 1 - l]) break;
				if (l > 0) p->cigar[k-1] -= l << 4, p->cigar[k+1] += l << 4, qoff -= l, toff -= l;
				if (l == prev_len) to_shrink = 1;
			}
			if (op == 1) qoff += len;
			else toff += len;
		} else if (op == 3) toff += len;
	}
	assert(qoff == r->qe - r->qs && toff == r->re - r->rs);
	for (k = 0; k + 2 < p->n_cigar; ++k) { /* runs like 5I6D7I become one I and one D */
		if ((p->cigar[k] & 0xf) > 0 && (p->cigar[k] & 0xf) + (p->cigar[k+1] & 0xf) == 3) {
			uint32_t l, s[3] = {0, 0, 0};
			for (l = k; l < p->n_cigar; ++l) {
				const uint32_t op = p->cigar[l] & 0xf;
				if (op == 1 || op == 2 || p->cigar[l] >> 4 == 0) s[op] += p->cigar[l] >> 4;
				else break;
			}
			if (s[1] > 0 && s[2] > 0 && l - k > 2) {
				p->cigar[k] = s[1] << 4 | 1;
				p->cigar[k+1] = s[2] << 4 | 2;
				for (k += 2; k < l; ++k) p->cigar[k] &= 0xf;
				to_shrink = 1;
			}
			k = l;
		}
	}
	if (to_shrink) {
		int32_t l = 0;
		for (k = 0; k < p->n_cigar; ++k)
			if (p->cigar[k] >> 4 != 0) p->cigar[l++] = p->cigar[k];
		p->n_cigar = l;
		for (k = l = 0; k < p->n_cigar; ++k)
			if (k == p->n_cigar - 1 || (p->cigar[k] & 0xf) != (p->cigar[k+1] & 0xf)) p->cigar[l++] = p->cigar[k];
			else p->cigar[k+1] += p->cigar[k] >> 4 << 4;
		p->n_cigar = l;
	}
	if ((p->cigar[0] & 0xf) == 1 || (p->cigar[0] & 0xf) == 2) {
		const int32_t l = p->cigar[0] >> 4;
		if ((p->cigar[0] & 0xf) == 1) {
			if (r->rev) r->qe -= l;
			else r->qs += l;
			*qshift = l;
		} else r->rs += l, *tshift = l;
		--p->n_cigar;
		memmove(p->cigar, p->cigar + 1, (size_t)p->n_cigar * 4);
	}
}

MM_FN static void cigar_to_eqx(mm_reg1_t *r, const uint8_t *qseq, const uint8_t *tseq)
{ /* align.c:169-238: M -> =/X */
	uint32_t n_EQX = 0, k, l, m, cap, toff = 0, qoff = 0, n_M = 0;
	mm_extra_t *p;
	if (r->p == 0) return;
	for (k = 0; k < r->p->n_cigar; ++k) {
		uint32_t op = r->p->cigar[k] & 0xf, len = r->p->cigar[k] >> 4;
		if (op == 0) {
			while (len > 0) {
				for (l = 0; l < len && qseq[qoff + l] == tseq[toff + l]; ++l) {}
				if (l > 0) { ++n_EQX; len -= l; toff += l; qoff += l; }
				for (l = 0; l < len && qseq[qoff + l] != tseq[toff + l]; ++l) {}
				if (l > 0) { ++n_EQX; len -= l; toff += l; qoff += l; }
			}
			++n_M;
		} else if (op == 1) qoff += len;
		else if (op == 2 || op == 3) toff += len;
	}
	if (n_EQX == n_M) { /* every M is a pure match run: rewrite in place */
		for (k = 0; k < r->p->n_cigar; ++k)
			if ((r->p->cigar[k] & 0xf) == 0) r->p->cigar[k] = r->p->cigar[k] >> 4 << 4 | 7;
		return;
	}
	cap = r->p->n_cigar + (n_EQX - n_M) + sizeof(mm_extra_t);
	cap = mm_roundup32(cap);
	p = (mm_extra_t*)calloc(cap, 4);
	memcpy(p, r->p, sizeof(mm_extra_t));
	p->capacity = cap;
	toff = qoff = m = 0;
	for (k = 0; k < r->p->n_cigar; ++k) {
		uint32_t op = r->p->cigar[k] & 0xf, len = r->p->cigar[k] >> 4;
		if (op == 0) {
			while (len > 0) {
				for (l = 0; l < len && qseq[qoff + l] == tseq[toff + l]; ++l) {}
				if (l > 0) p->cigar[m++] = l << 4 | 7;
				len -= l, toff += l, qoff += l;
				for (l = 0; l < len && qseq[qoff + l] != tseq[toff + l]; ++l) {}
				if (l > 0) p->cigar[m++] = l << 4 | 8;
				len -= l, toff += l, qoff += l;
			}
			continue;
		} else if (op == 1) qoff += len;
		else if (op == 2 || op == 3) toff += len;
		p->cigar[m++] = r->p->cigar[k];
	}
	p->n_cigar = m;
	free(r->p);
	r->p = p;
}

MM_FN static void update_extra(mm_reg1_t *r, const uint8_t *qseq, const uint8_t *tseq, const int8_t *mat, int8_t q, int8_t e, int is_eqx)
{ /* align.c:240-286: final CIGAR clean-up, then blen/mlen/n_ambi and the local-style dp_max */
	uint32_t k, l;
	int32_t s = 0, max = 0, qshift, tshift, toff = 0, qoff = 0;
	mm_extra_t *p = r->p;
	if (p == 0) return;
	fix_cigar(r, qseq, tseq, &qshift, &tshift);
	qseq += qshift, tseq += tshift;
	r->blen = r->mlen = 0;
	for (k = 0; k < p->n_cigar; ++k) {
		const uint32_t op = p->cigar[k] & 0xf, len = p->cigar[k] >> 4;
		if (op == 0) {
			int n_ambi = 0, n_diff = 0;
			for (l = 0; l < len; ++l) {
				const int cq = qseq[qoff + l], ct = tseq[toff + l];
				if (ct > 3 || cq > 3) ++n_ambi;
				else if (ct != cq) ++n_diff;
				s += mat[ct * 5 + cq];
				if (s < 0) s = 0;
				else max = max > s ? max : s;
			}
			r->blen += len - n_ambi, r->mlen += len - (n_ambi + n_diff), p->n_ambi += n_ambi;
			toff += len, qoff += len;
		} else if (op == 1 || op == 2) {
			const uint8_t *sq = op == 1 ? qseq + qoff : tseq + toff;
			int n_ambi = 0;
			for (l = 0; l < len; ++l) if (sq[l] > 3) ++n_ambi;
			r->blen += len - n_ambi, p->n_ambi += n_ambi;
			s -= q + e * len;
			if (s < 0) s = 0;
			if (op == 1) qoff += len; else toff += len;
		} else if (op == 3) toff += len;
	}
	p->dp_max = max;
	assert(qoff == r->qe - r->qs && toff == r->re - r->rs);
	if (is_eqx) cigar_to_eqx(r, qseq, tseq);
}

/* ---- z-drop re-test on a finished gap fill (align.c:32-89) */

MM_FN static int test_zdrop(const walk_t *w, const uint8_t *qseq, const uint8_t *tseq, uint32_t n_cigar, const uint32_t *cigar)
{
	const mm_mapopt_t *opt = w->opt;
	uint32_t k;
	int32_t score = 0, max = INT32_MIN, max_i = -1, max_j = -1, i = 0, j = 0, max_zdrop = 0;
	int pos[2][2] = {{-1, -1}, {-1, -1}}, q_len, t_len;
#define TRACK(ii, jj) do { \
		if (score < max) { \
			const int li = (ii) - max_i, lj = (jj) - max_j, diff = li > lj ? li - lj : lj - li, z = max - score - diff * opt->e; \
			if (z > max_zdrop) max_zdrop = z, pos[0][0] = max_i, pos[0][1] = (ii) + 1, pos[1][0] = max_j, pos[1][1] = (jj) + 1; \
		} else max = score, max_i = (ii), max_j = (jj); \
	} while (0)
	for (k = 0; k < n_cigar; ++k) {
		const uint32_t op = cigar[k] & 0xf, len = cigar[k] >> 4;
		uint32_t l;
		if (op == 0) {
			for (l = 0; l < len; ++l) {
				score += w->mat[tseq[i + l] * 5 + qseq[j + l]];
				TRACK(i + (int32_t)l, j + (int32_t)l);
			}
			i += len, j += len;
		} else if (op == 1 || op == 2 || op == 3) {
			score -= opt->q + opt->e * len;
			if (op == 1) j += len;
			else i += len;
			TRACK(i, j);
		}
	}
#undef TRACK
	q_len = pos[1][1] - pos[1][0], t_len = pos[0][1] - pos[0][0];
	if (!(opt->flag & (MM_F_SPLICE|MM_F_SR|MM_F_FOR_ONLY|MM_F_REV_ONLY)) && max_zdrop > opt->zdrop_inv && q_len < opt->max_gap && t_len < opt->max_gap) {
		uint8_t *qseq2 = (uint8_t*)mm_amalloc(q_len > 0 ? q_len : 1);
		int q_off, t_off;
		for (i = 0; i < q_len; ++i) {
			const int c = qseq[pos[1][1] - i - 1];
			qseq2[i] = c >= 4 ? 4 : 3 - c;
		}
		score = mm_ll_i16(q_len, qseq2, t_len, tseq + pos[0][0], 5, w->mat, opt->q, opt->e, &q_off, &t_off);
		mm_afree(qseq2);
		if (score >= opt->min_chain_score * opt->a && score >= opt->min_dp_max) return 2; /* looks like an inversion */
	}
	return max_zdrop > opt->zdrop ? 1 : 0;
}

/* ---- which anchors bound the DP (align.c:341-521) */

MM_FN static void adjust_minier(const mm_idx_t *mi, uint8_t *const qseq0[2], const mm128_t *a, int32_t *r, int32_t *q)
{ /* align.c:341-365 */
	if (mi->flag & MM_I_HPC) {
		const uint8_t *qseq = qseq0[a->x >> 63];
		const uint32_t rid = a->x << 1 >> 33;
		const int64_t off0 = mi->seq[rid].offset, off = off0 + (int32_t)a->x;
		int64_t i;
		int c;
		*q = (int32_t)a->y;
		for (i = *q - 1, c = qseq[*q]; i > 0; --i) if (qseq[i] != c) break;
		*q = (int32_t)i + 1;
		c = mm_seq4_get(mi->S, off);
		for (i = off - 1; i >= off0; --i) if ((int)mm_seq4_get(mi->S, i) != c) break;
		*r = (int32_t)a->x + 1 - (int)(off - i);
	} else {
		*r = (int32_t)a->x - (mi->k >> 1);
		*q = (int32_t)a->y - (mi->k >> 1);
	}
}

#define GAP_AT(a, i) (((int32_t)(a)[i].y - (int32_t)(a)[(i) - 1].y) - ((int32_t)(a)[i].x - (int32_t)(a)[(i) - 1].x))

MM_FN static int *long_gaps(int as1, int cnt1, const mm128_t *a, int min_gap, int *n_)
{ /* align.c:367-384: anchor indices after which the diagonal moves by more than min_gap */
	int i, n = 0, *K;
	*n_ = 0;
	for (i = 1; i < cnt1; ++i) { const int gap = GAP_AT(a + as1, i); if (gap < -min_gap || gap > min_gap) ++n; }
	if (n <= 1) return 0;
	K = (int*)mm_amalloc((size_t)n * sizeof(int));
	for (i = 1, n = 0; i < cnt1; ++i) { const int gap = GAP_AT(a + as1, i); if (gap < -min_gap || gap > min_gap) K[n++] = i; }
	*n_ = n;
	return K;
}

MM_FN static void filter_bad_seeds(int as1, int cnt1, mm128_t *a, int min_gap, int diff_thres, int max_ext_len, int max_ext_cnt)
{ /* align.c:386-421: ignore seeds between an insertion and a compensating deletion */
	int max_st = -1, max_en = -1, n, i, k, max = 0, *K;
	mm128_t *b = a + as1;
	K = long_gaps(as1, cnt1, a, min_gap, &n);
	if (K == 0) return;
	for (k = 0;; ++k) {
		int gap, l, n_ins = 0, n_del = 0, qs, rs, max_diff = 0, max_diff_l = -1;
		if (k == n || k >= max_en) {
			if (max_en > 0) for (i = K[max_st]; i < K[max_en]; ++i) b[i].y |= MM_SEED_IGNORE;
			max = 0, max_st = max_en = -1;
			if (k == n) break;
		}
		i = K[k];
		gap = GAP_AT(b, i);
		if (gap > 0) n_ins += gap; else n_del += -gap;
		qs = (int32_t)b[i-1].y, rs = (int32_t)b[i-1].x;
		for (l = k + 1; l < n && l <= k + max_ext_cnt; ++l) {
			const int j = K[l];
			int diff;
			if ((int32_t)b[j].y - qs > max_ext_len || (int32_t)b[j].x - rs > max_ext_len) break;
			gap = GAP_AT(b, j);
			if (gap > 0) n_ins += gap; else n_del += -gap;
			diff = n_ins + n_del - abs(n_ins - n_del);
			if (max_diff < diff) max_diff = diff, max_diff_l = l;
		}
		if (max_diff > diff_thres && max_diff > max) max = max_diff, max_st = k, max_en = max_diff_l;
	}
	mm_afree(K);
}

MM_FN static void filter_bad_seeds_alt(int as1, int cnt1, mm128_t *a, int min_gap, int max_ext)
{ /* align.c:423-457: merge nearby long gaps into one long-join DP */
	int n, k, *K;
	mm128_t *b = a + as1;
	K = long_gaps(as1, cnt1, a, min_gap, &n);
	if (K == 0) return;
	for (k = 0; k < n;) {
		const int i = K[k];
		int l, gap1 = GAP_AT(b, i), re1 = (int32_t)b[i].x, qe1 = (int32_t)b[i].y;
		gap1 = gap1 > 0 ? gap1 : -gap1;
		for (l = k + 1; l < n; ++l) {
			const int j = K[l];
			int gap2, q_span_pre, rs2, qs2, m;
			if ((int32_t)b[j].y - qe1 > max_ext || (int32_t)b[j].x - re1 > max_ext) break;
			gap2 = GAP_AT(b, j);
			q_span_pre = b[j-1].y >> 32 & 0xff;
			rs2 = (int32_t)b[j-1].x + q_span_pre, qs2 = (int32_t)b[j-1].y + q_span_pre;
			m = rs2 - re1 < qs2 - qe1 ? rs2 - re1 : qs2 - qe1;
			gap2 = gap2 > 0 ? gap2 : -gap2;
			if (m > gap1 + gap2) break;
			re1 = (int32_t)b[j].x, qe1 = (int32_t)b[j].y, gap1 = gap2;
		}
		if (l > k + 1) {
			const int end = K[l - 1];
			int j;
			for (j = K[k]; j < end; ++j) b[j].y |= MM_SEED_IGNORE;
			b[end].y |= MM_SEED_LONG_JOIN;
		}
		k = l;
	}
	mm_afree(K);
}

MM_FN static void fix_bad_ends(const mm_reg1_t *r, const mm128_t *a, int bw, int min_match, int32_t *as, int32_t *cnt)
{ /* align.c:459-493: trim end anchors that sit off the main diagonal */
	int32_t i, l, m;
	*as = r->as, *cnt = r->cnt;
	if (r->cnt < 3) return;
	m = l = a[r->as].y >> 32 & 0xff;
	for (i = r->as + 1; i < r->as + r->cnt - 1; ++i) {
		const int32_t q_span = a[i].y >> 32 & 0xff;
		int32_t lq, lr, min, max;
		if (a[i].y & MM_SEED_LONG_JOIN) break;
		lr = (int32_t)a[i].x - (int32_t)a[i-1].x, lq = (int32_t)a[i].y - (int32_t)a[i-1].y;
		min = lr < lq ? lr : lq, max = lr > lq ? lr : lq;
		if (max - min > l >> 1) *as = i;
		l += min;
		m += min < q_span ? min : q_span;
		if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r->mlen >> 1) break;
	}
	*cnt = r->as + r->cnt - *as;
	m = l = a[r->as + r->cnt - 1].y >> 32 & 0xff;
	for (i = r->as + r->cnt - 2; i > *as; --i) {
		const int32_t q_span = a[i+1].y >> 32 & 0xff;
		int32_t lq, lr, min, max;
		if (a[i+1].y & MM_SEED_LONG_JOIN) break;
		lr = (int32_t)a[i+1].x - (int32_t)a[i].x, lq = (int32_t)a[i+1].y - (int32_t)a[i].y;
		min = lr < lq ? lr : lq, max = lr > lq ? lr : lq;
		if (max - min > l >> 1) *cnt = i + 1 - *as;
		l += min;
		m += min < q_span ? min : q_span;
		if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r->mlen >> 1) break;
	}
}

MM_FN static void max_stretch(const mm_reg1_t *r, const mm128_t *a, int32_t *as, int32_t *cnt)
{ /* align.c:495-521: short reads keep only the best run of anchors on one diagonal */
	int32_t i, score, max_score = -1, len, max_i = -1, max_len = 0;
	*as = r->as, *cnt = r->cnt;
	if (r->cnt < 2) return;
	score = a[r->as].y >> 32 & 0xff, len = 1;
	for (i = r->as + 1; i < r->as + r->cnt; ++i) {
		const int32_t q_span = a[i].y >> 32 & 0xff;
		const int32_t lr = (int32_t)a[i].x - (int32_t)a[i-1].x, lq = (int32_t)a[i].y - (int32_t)a[i-1].y;
		if (lq == lr) score += lq < q_span ? lq : q_span, ++len;
		else {
			if (score > max_score) max_score = score, max_len = len, max_i = i - len;
			score = q_span, len = 1;
		}
	}
	if (score > max_score) max_score = score, max_len = len, max_i = i - len;
	*as = max_i, *cnt = max_len;
}

/* ---- one region (align.c:565-788) */

MM_FN static void align1(walk_t *w, mm_reg1_t *r, mm_reg1_t *r2)
{
	const mm_mapopt_t *opt = w->opt;
	const mm_idx_t *mi = w->mi;
	mm_alnseg_t *sg = w->seg;
	mm128_t *a = sg->a;
	uint8_t *const *qseq0 = sg->qseq0;
	const int qlen = sg->qlen, n_a = sg->n_a;
	const int is_sr = !!(opt->flag & MM_F_SR);
	const int32_t rid = a[r->as].x << 1 >> 33, rev = a[r->as].x >> 63;
	const int32_t ref_len = (int32_t)mi->seq[rid].len;
	int32_t as1, cnt1, i, l, bw, dropped = 0, rs0, re0, qs0, qe0, rs, re, qs, qe, rs1, qs1, re1, qe1;
	uint8_t *tseq;
	ez_t ez;

	r2->cnt = 0;
	if (r->cnt == 0) return;
	bw = (int)(opt->bw * 1.5 + 1.);

	if (is_sr && !(mi->flag & MM_I_HPC)) {
		max_stretch(r, a, &as1, &cnt1);
		rs = (int32_t)a[as1].x + 1 - (int32_t)(a[as1].y >> 32 & 0xff);
		qs = (int32_t)a[as1].y + 1 - (int32_t)(a[as1].y >> 32 & 0xff);
		re = (int32_t)a[as1 + cnt1 - 1].x + 1;
		qe = (int32_t)a[as1 + cnt1 - 1].y + 1;
	} else {
		if (!(opt->flag & MM_F_NO_END_FLT)) fix_bad_ends(r, a, opt->bw, opt->min_chain_score * 2, &as1, &cnt1);
		else as1 = r->as, cnt1 = r->cnt;
		filter_bad_seeds(as1, cnt1, a, 10, 40, opt->max_gap >> 1, 10);
		filter_bad_seeds_alt(as1, cnt1, a, 30, opt->max_gap >> 1);
		adjust_minier(mi, qseq0, &a[as1], &rs, &qs);
		adjust_minier(mi, qseq0, &a[as1 + cnt1 - 1], &re, &qe);
	}
	assert(cnt1 > 0);

	/* the window handed to the two extensions (align.c:613-684) */
	if (is_sr) {
		qs0 = 0, qe0 = qlen;
		l = qs;
		l += l * opt->a + opt->end_bonus > opt->q ? (l * opt->a + opt->end_bonus - opt->q) / opt->e : 0;
		rs0 = rs - l > 0 ? rs - l : 0;
		l = qlen - qe;
		l += l * opt->a + opt->end_bonus > opt->q ? (l * opt->a + opt->end_bonus - opt->q) / opt->e : 0;
		re0 = re + l < ref_len ? re + l : ref_len;
	} else {
		rs0 = (int32_t)a[r->as].x + 1 - (int32_t)(a[r->as].y >> 32 & 0xff);
		qs0 = (int32_t)a[r->as].y + 1 - (int32_t)(a[r->as].y >> 32 & 0xff);
		if (rs0 < 0) rs0 = 0;
		assert(qs0 >= 0);
		rs1 = qs1 = 0;
		for (i = r->as - 1, l = 0; i >= 0 && a[i].x >> 32 == a[r->as].x >> 32; --i) { /* seeds of neighbouring chains limit the reach */
			const int32_t x = (int32_t)a[i].x + 1 - (int32_t)(a[i].y >> 32 & 0xff);
			const int32_t y = (int32_t)a[i].y + 1 - (int32_t)(a[i].y >> 32 & 0xff);
			if (x < rs0 && y < qs0) {
				if (++l > opt->min_cnt) {
					l = rs0 - x > qs0 - y ? rs0 - x : qs0 - y;
					rs1 = rs0 - l, qs1 = qs0 - l;
					if (rs1 < 0) rs1 = 0;
					break;
				}
			}
		}
		if (qs > 0 && rs > 0) {
			l = qs < opt->max_gap ? qs : opt->max_gap;
			qs1 = qs1 > qs - l ? qs1 : qs - l;
			qs0 = qs0 < qs1 ? qs0 : qs1;
			l += l * opt->a > opt->q ? (l * opt->a - opt->q) / opt->e : 0;
			l = l < opt->max_gap ? l : opt->max_gap;
			l = l < rs ? l : rs;
			rs1 = rs1 > rs - l ? rs1 : rs - l;
			rs0 = rs0 < rs1 ? rs0 : rs1;
			rs0 = rs0 < rs ? rs0 : rs;
		} else rs0 = rs, qs0 = qs;
		re0 = (int32_t)a[r->as + r->cnt - 1].x + 1;
		qe0 = (int32_t)a[r->as + r->cnt - 1].y + 1;
		re1 = ref_len, qe1 = qlen;
		for (i = r->as + r->cnt, l = 0; i < n_a && a[i].x >> 32 == a[r->as].x >> 32; ++i) {
			const int32_t x = (int32_t)a[i].x + 1, y = (int32_t)a[i].y + 1;
			if (x > re0 && y > qe0) {
				if (++l > opt->min_cnt) {
					l = x - re0 > y - qe0 ? x - re0 : y - qe0;
					re1 = re0 + l, qe1 = qe0 + l;
					break;
				}
			}
		}
		if (qe < qlen && re < ref_len) {
			l = qlen - qe < opt->max_gap ? qlen - qe : opt->max_gap;
			qe1 = qe1 < qe + l ? qe1 : qe + l;
			qe0 = qe0 > qe1 ? qe0 : qe1;
			l += l * opt->a > opt->q ? (l * opt->a - opt->q) / opt->e : 0;
			l = l < opt->max_gap ? l : opt->max_gap;
			l = l < ref_len - re ? l : ref_len - re;
			re1 = re1 < re + l ? re1 : re + l;
			re0 = re0 > re1 ? re0 : re1;
		} else re0 = re, qe0 = qe;
	}
	if (a[r->as].y & MM_SEED_SELF) {
		int max_ext = r->qs > r->rs ? r->qs - r->rs : r->rs - r->qs;
		if (r->rs - rs0 > max_ext) rs0 = r->rs - max_ext;
		if (r->qs - qs0 > max_ext) qs0 = r->qs - max_ext;
		max_ext = r->qe > r->re ? r->qe - r->re : r->re - r->qe;
		if (re0 - r->re > max_ext) re0 = r->re + max_ext;
		if (qe0 - r->qe > max_ext) qe0 = r->qe + max_ext;
	}
	assert(re0 > rs0);
	tseq = (uint8_t*)mm_amalloc((size_t)(re0 - rs0));

	if (qs > 0 && rs > 0) { /* left extension: both slices reversed, gaps right-aligned, CIGAR reversed (align.c:690-705) */
		dp_request(w, rev, qs0, qs - qs0, rid, rs0, rs - rs0, 1, bw, opt->end_bonus, r->split_inv ? opt->zdrop_inv : opt->zdrop,
		           KSW_EZ_EXTZ_ONLY|KSW_EZ_RIGHT|KSW_EZ_REV_CIGAR, &ez);
		if (ez.n_cigar > 0) {
			append_cigar(r, ez.n_cigar, ez.cigar);
			r->p->dp_score += ez.max;
		}
		rs1 = rs - (ez.reach_end ? ez.mqe_t + 1 : ez.max_t + 1);
		qs1 = qs - (ez.reach_end ? qs - qs0 : ez.max_q + 1);
	} else rs1 = rs, qs1 = qs;
	re1 = rs, qe1 = qs;
	assert(w->pending || (qs1 >= 0 && rs1 >= 0));

	for (i = is_sr ? cnt1 - 1 : 1; i < cnt1; ++i) { /* gap filling (align.c:709-758) */
		if ((a[as1 + i].y & (MM_SEED_IGNORE|MM_SEED_TANDEM)) && i != cnt1 - 1) continue;
		if (is_sr && !(mi->flag & MM_I_HPC)) re = (int32_t)a[as1 + i].x + 1, qe = (int32_t)a[as1 + i].y + 1;
		else adjust_minier(mi, qseq0, &a[as1 + i], &re, &qe);
		re1 = re, qe1 = qe;
		if (i == cnt1 - 1 || (a[as1 + i].y & MM_SEED_LONG_JOIN) || (qe - qs >= opt->min_ksw_len && re - rs >= opt->min_ksw_len)) {
			int j, bw1 = bw, zdrop_code;
			const uint8_t *qseq = &qseq0[rev][qs];
			uint32_t one_op;
			if (a[as1 + i].y & MM_SEED_LONG_JOIN) bw1 = qe - qs > re - rs ? qe - qs : re - rs;
			if (is_sr) { /* short reads: no DP between the ends of the stretch; N scores +e2 (align.c:724-731) */
				mm_fillmemo_t *fm = 0;
				int q;
				assert(qe - qs == re - rs);
				ez_reset(&ez);
				for (q = 0; q < sg->n_fill; ++q) /* a region is walked twice (request DP, consume DP): score its stretch once */
					if (sg->fill[q].rev == rev && sg->fill[q].rid == rid && sg->fill[q].rs == rs && sg->fill[q].qs == qs && sg->fill[q].len == qe - qs) { fm = &sg->fill[q]; break; }
				if (fm == 0) {
					mm_idx_getseq(mi, rid, rs, re, tseq);
					for (j = 0, ez.score = 0; j < qe - qs; ++j) {
						if (qseq[j] >= 4 || tseq[j] >= 4) ez.score += opt->e2;
						else ez.score += qseq[j] == tseq[j] ? opt->a : -opt->b;
					}
					one_op = (uint32_t)(qe - qs) << 4 | 0;
					zdrop_code = test_zdrop(w, qseq, tseq, 1, &one_op);
					if (sg->n_fill == sg->m_fill) {
						const size_t old_b = (size_t)sg->m_fill * sizeof(mm_fillmemo_t);
						sg->m_fill = sg->m_fill ? sg->m_fill << 1 : 4;
						sg->fill = (mm_fillmemo_t*)mm_arealloc(sg->fill, old_b, (size_t)sg->m_fill * sizeof(mm_fillmemo_t));
					}
					fm = &sg->fill[sg->n_fill++];
					fm->rev = rev, fm->rid = rid, fm->rs = rs, fm->qs = qs, fm->len = qe - qs, fm->score = ez.score, fm->zdrop_code = zdrop_code;
				}
				ez.score = fm->score, zdrop_code = fm->zdrop_code;
				one_op = (uint32_t)(qe - qs) << 4 | 0;
				ez.cigar = &one_op, ez.n_cigar = 1;
			} else {
				mm_idx_getseq(mi, rid, rs, re, tseq);
				dp_request(w, rev, qs, qe - qs, rid, rs, re - rs, 0, bw1, -1, opt->zdrop, KSW_EZ_APPROX_MAX, &ez);
				zdrop_code = test_zdrop(w, qseq, tseq, ez.n_cigar, ez.cigar);
			}
			if (zdrop_code != 0) /* second, exact pass */
				dp_request(w, rev, qs, qe - qs, rid, rs, re - rs, 0, bw1, -1, zdrop_code == 2 ? opt->zdrop_inv : opt->zdrop, 0, &ez);
			if (ez.n_cigar > 0) append_cigar(r, ez.n_cigar, ez.cigar);
			if (ez.zdropped) { /* cut the region here; the rest becomes r2 (align.c:741-754) */
				for (j = i - 1; j >= 0; --j)
					if ((int32_t)a[as1 + j].x <= rs + ez.max_t) break;
				dropped = 1;
				if (j < 0) j = 0;
				r->p->dp_score += ez.max;
				re1 = rs + (ez.max_t + 1);
				qe1 = qs + (ez.max_q + 1);
				if (cnt1 - (j + 1) >= opt->min_cnt) {
					mm_split_reg(r, r2, as1 + j + 1 - r->as, qlen, a);
					if (zdrop_code == 2) r2->split_inv = 1;
				}
				break;
			} else if (r->p) r->p->dp_score += ez.score;
			rs = re, qs = qe;
		}
	}

	if (!dropped && qe < qe0 && re < re0) { /* right extension (align.c:760-771) */
		dp_request(w, rev, qe, qe0 - qe, rid, re, re0 - re, 0, bw, opt->end_bonus, opt->zdrop, KSW_EZ_EXTZ_ONLY, &ez);
		if (ez.n_cigar > 0) {
			append_cigar(r, ez.n_cigar, ez.cigar);
			r->p->dp_score += ez.max;
		}
		re1 = re + (ez.reach_end ? ez.mqe_t + 1 : ez.max_t + 1);
		qe1 = qe + (ez.reach_end ? qe0 - qe : ez.max_q + 1);
	}
	if (w->pending) { mm_afree(tseq); return; } /* the caller throws this attempt away */
	assert(qe1 <= qlen);

	r->rs = rs1, r->re = re1;
	if (rev) r->qs = qlen - qe1, r->qe = qlen - qs1;
	else r->qs = qs1, r->qe = qe1;
	assert(re1 - rs1 <= re0 - rs0);
	if (r->p) {
		mm_idx_getseq(mi, rid, rs1, re1, tseq);
		update_extra(r, &qseq0[r->rev][qs1], tseq, w->mat, opt->q, opt->e, opt->flag & MM_F_EQX);
		if (rev && r->p->trans_strand) r->p->trans_strand ^= 3;
	}
	mm_afree(tseq);
}

/* ---- inversion rescue between two pieces of a z-drop split (align.c:790-845) */

MM_FN static int align1_inv(walk_t *w, const mm_reg1_t *r1, const mm_reg1_t *r2, mm_reg1_t *r_inv)
{
	const mm_mapopt_t *opt = w->opt;
	const mm_idx_t *mi = w->mi;
	mm_alnseg_t *sg = w->seg;
	const int qlen = sg->qlen;
	int tl, ql, score, ret = 0, q_off, t_off, q_rev, q_base;
	uint8_t *tseq, *qrev, *trev;
	const uint8_t *qseq;
	ez_t ez;

	memset(r_inv, 0, sizeof(mm_reg1_t));
	if (!(r1->split & 1) || !(r2->split & 2)) return 0;
	if (r1->id != r1->parent && r1->parent != MM_PARENT_TMP_PRI) return 0;
	if (r2->id != r2->parent && r2->parent != MM_PARENT_TMP_PRI) return 0;
	if (r1->rid != r2->rid || r1->rev != r2->rev) return 0;
	ql = r1->rev ? r1->qs - r2->qe : r2->qs - r1->qe;
	tl = r2->rs - r1->re;
	if (ql < opt->min_chain_score || ql > opt->max_gap) return 0;
	if (tl < opt->min_chain_score || tl > opt->max_gap) return 0;

	tseq = (uint8_t*)mm_amalloc((size_t)tl * 2 + (size_t)ql);
	trev = tseq + tl, qrev = trev + tl;
	mm_idx_getseq(mi, r1->rid, r1->re, r2->rs, tseq);
	q_rev = r1->rev ? 0 : 1, q_base = r1->rev ? r2->qe : qlen - r2->qs; /* the opposite strand of the gap (align.c:810) */
	qseq = &sg->qseq0[q_rev][q_base];
	for (int i = 0; i < ql; ++i) qrev[i] = qseq[ql - 1 - i];
	for (int i = 0; i < tl; ++i) trev[i] = tseq[tl - 1 - i];
	score = mm_ll_i16(ql, qrev, tl, trev, 5, w->mat, opt->q, opt->e, &q_off, &t_off);
	if (score < opt->min_dp_max) goto done;
	q_off = ql - (q_off + 1), t_off = tl - (t_off + 1);
	if (!dp_request(w, q_rev, q_base + q_off, ql - q_off, r1->rid, r1->re + t_off, tl - t_off, 0, (int)(opt->bw * 1.5), -1, opt->zdrop, KSW_EZ_EXTZ_ONLY, &ez)) {
		ret = -1; /* pending */
		goto done;
	}
	if (ez.n_cigar == 0) goto done;
	append_cigar(r_inv, ez.n_cigar, ez.cigar);
	r_inv->p->dp_score = ez.max;
	r_inv->id = -1, r_inv->parent = MM_PARENT_UNSET, r_inv->inv = 1;
	r_inv->rev = !r1->rev, r_inv->rid = r1->rid, r_inv->div = -1.0f;
	if (r_inv->rev == 0) r_inv->qs = r2->qe + q_off, r_inv->qe = r_inv->qs + ez.max_q + 1;
	else r_inv->qe = r2->qs - q_off, r_inv->qs = r_inv->qe - (ez.max_q + 1);
	r_inv->rs = r1->re + t_off, r_inv->re = r_inv->rs + ez.max_t + 1;
	update_extra(r_inv, &qseq[q_off], &tseq[t_off], w->mat, opt->q, opt->e, opt->flag & MM_F_EQX);
	ret = 1;
done:
	mm_afree(tseq);
	return ret;
}

/* ---- the resumable skeleton (align.c:857-913) */

MM_FN void mm_aln_begin(mm_alnseg_t *s, int seq_id, int qlen, const char *qstr, int n_regs, mm_reg1_t *regs, mm128_t *a)
{
	memset(s, 0, sizeof(*s));
	s->seq_id = seq_id, s->qlen = qlen, s->qstr = qstr, s->n_regs = n_regs, s->regs = regs, s->a = a;
}

MM_FN static mm_reg1_t *insert_reg(const mm_reg1_t *r, int i, int *n_regs, mm_reg1_t *regs)
{ /* align.c:847-855 */
	regs = (mm_reg1_t*)realloc(regs, (size_t)(*n_regs + 1) * sizeof(mm_reg1_t));
	if (i + 1 != *n_regs) memmove(&regs[i + 2], &regs[i + 1], sizeof(mm_reg1_t) * (size_t)(*n_regs - i - 1));
	regs[i + 1] = *r;
	++*n_regs;
	return regs;
}

MM_FN int mm_aln_step(mm_alnseg_t *s, const mm_mapopt_t *opt, const mm_idx_t *mi)
{
	walk_t w;
	int i;
	if (s->finished) return 1;
	if (opt->flag & MM_F_SPLICE) { fprintf(stderr, "[ERROR] spliced alignment is outside the scope of this build\n"); exit(1); }
	w.seg = s, w.opt = opt, w.mi = mi, w.pending = 0;
	gen_simple_mat(w.mat, opt->a, opt->b, opt->sc_ambi);
	if (!s->started) { /* encode both strands of the query, compact the anchors (align.c:864-873) */
		s->qseq0[0] = (uint8_t*)mm_amalloc((size_t)(s->qlen > 0 ? s->qlen : 1) * 2);
		s->qseq0[1] = s->qseq0[0] + s->qlen;
		for (i = 0; i < s->qlen; ++i) {
#ifdef MM_DEVICE_BUILD
			s->qseq0[0][i] = (uint8_t)mm_seq4_get(s->q4, s->q4_off + i); /* encoded by k_encode_reads with the same table */
#else
			s->qseq0[0][i] = seq_nt4_table[(uint8_t)s->qstr[i]];
#endif
			s->qseq0[1][s->qlen - 1 - i] = s->qseq0[0][i] < 4 ? 3 - s->qseq0[0][i] : 4;
		}
		s->n_a = mm_squeeze_a(s->n_regs, s->regs, s->a);
		s->started = 1, s->i = 0;
	}
	while (s->i < s->n_regs) {
		i = s->i;
		if (!s->inv_wait) {
			mm_reg1_t saved = s->regs[i], r2;
			w.pending = 0;
			align1(&w, &s->regs[i], &r2);
			if (w.pending) { /* discard the partial attempt; it is replayed once its DP jobs are back */
				int j;
				if (s->regs[i].p != saved.p) free(s->regs[i].p);
				s->regs[i] = saved;
				/* regions do not depend on each other's DP results: request the jobs of all later regions now,
				 * so that a segment normally needs two rounds (plan, consume) however many regions it has */
				for (j = (s->planned > i ? s->planned : i) + 1; j < s->n_regs; ++j) {
					mm_reg1_t tmp = s->regs[j], t2;
					w.pending = 0;
					align1(&w, &tmp, &t2);
					if (tmp.p != s->regs[j].p) free(tmp.p);
				}
				s->planned = s->n_regs - 1;
				return 0;
			}
			if (r2.cnt > 0) {
				s->regs = insert_reg(&r2, i, &s->n_regs, s->regs);
				if (s->planned > i) ++s->planned;
			}
		}
		s->inv_wait = 0;
		if (i > 0 && s->regs[i].split_inv) {
			mm_reg1_t r_inv;
			int rc;
			w.pending = 0;
			rc = align1_inv(&w, &s->regs[i-1], &s->regs[i], &r_inv);
			if (rc < 0) { s->inv_wait = 1; return 0; } /* region i is final; only the inversion DP is outstanding */
			if (rc > 0) { s->regs = insert_reg(&r_inv, i, &s->n_regs, s->regs); ++i; if (s->planned >= i) ++s->planned; }
		}
		s->i = i + 1;
	}
	mm_afree(s->qseq0[0]); s->qseq0[0] = s->qseq0[1] = 0;
	mm_filter_regs(opt, s->qlen, &s->n_regs, s->regs);
	mm_hit_sort(&s->n_regs, s->regs);
	s->finished = 1;
	return 1;
}

MM_FN void mm_aln_end(mm_alnseg_t *s)
{
	int i;
	for (i = 0; i < s->cache.n; ++i) mm_afree(s->cache.a[i].cigar);
	mm_afree(s->cache.a);
	mm_afree(s->qseq0[0]);
	memset(&s->cache, 0, sizeof(s->cache));
	s->qseq0[0] = s->qseq0[1] = 0;
}
