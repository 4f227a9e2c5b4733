/* hits.c -- chains -> hits: region records, primary/secondary hierarchy, filters, long-join,
 * per-mate split, MAPQ, pairing and the divergence estimate.  Host side on purpose: this is
 * O(#chains) float/logf bookkeeping whose results must match glibc bit for bit.
 * Reference: hit.c, pe.c, esterr.c of src/minimap2-master_remapping (cited per function). */
#include <math.h>
#include <stdio.h>
#include "mm2b_priv.h"

#define SPAN(an) ((int32_t)((an).y >> 32 & 0xff))
#define RPOS(an) ((int32_t)(an).x)
#define QPOS(an) ((int32_t)(an).y)

MM_FN static inline uint32_t wang_hash(uint32_t key) /* khash.h:400-409 */
{
	key += ~(key << 15); key ^= (key >> 10); key += (key << 3);
	key ^= (key >> 6);   key += ~(key << 11); key ^= (key >> 16);
	return key;
}

#ifndef MM_DEVICE_BUILD
MM_FN uint32_t mm_frag_hash(const char *qname, int qlen_sum, int seed)
{ /* the per-fragment salt of map.c:291-293: X31 string hash of the name, mixed with length and seed */
	uint32_t h = 0;
	if (qname) {
		const char *s = qname;
		h = (uint32_t)*s;
		if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)*s;
	}
	h ^= wang_hash((uint32_t)qlen_sum) + wang_hash((uint32_t)seed);
	return wang_hash(h);
}
#endif

MM_FN static inline uint64_t mix64(uint64_t key) /* hit.c:40-50 (unmasked variant of the sketch hash) */
{
	key = ~key + (key << 21);
	key ^= key >> 24;
	key = key + (key << 3) + (key << 8);
	key ^= key >> 14;
	key = key + (key << 2) + (key << 4);
	key ^= key >> 28;
	key = key + (key << 31);
	return key;
}

MM_FN static void reg_set_coor(mm_reg1_t *r, int32_t qlen, const mm128_t *a)
{ /* hit.c:8-38: coordinates and the fuzzy match/block lengths from the chain's anchors */
	const mm128_t *first = &a[r->as], *last = &a[r->as + r->cnt - 1];
	const int32_t q_span = SPAN(*first);
	int i;
	r->rev = first->x >> 63;
	r->rid = first->x << 1 >> 33;
	r->rs = RPOS(*first) + 1 > q_span ? RPOS(*first) + 1 - q_span : 0;
	r->re = RPOS(*last) + 1;
	if (!r->rev) r->qs = QPOS(*first) + 1 - q_span, r->qe = QPOS(*last) + 1;
	else r->qs = qlen - (QPOS(*last) + 1), r->qe = qlen - (QPOS(*first) + 1 - q_span);
	r->mlen = r->blen = 0;
	if (r->cnt <= 0) return;
	r->mlen = r->blen = q_span;
	for (i = r->as + 1; i < r->as + r->cnt; ++i) {
		const int span = SPAN(a[i]), tl = RPOS(a[i]) - RPOS(a[i-1]), ql = QPOS(a[i]) - QPOS(a[i-1]);
		r->blen += tl > ql ? tl : ql;
		r->mlen += tl > span && ql > span ? span : tl < ql ? tl : ql;
	}
}

MM_FN mm_reg1_t *mm_gen_regs(uint32_t hash, int qlen, int n_u, uint64_t *u, mm128_t *a)
{ /* hit.c:52-88 */
	mm128_t *z;
	mm_reg1_t *r;
	int i, k;
	if (n_u == 0) return 0;
	z = (mm128_t*)mm_amalloc((size_t)n_u * 16);
	for (i = k = 0; i < n_u; ++i) { /* order: chain score, ties broken by a salted hash of the first anchor */
		const uint32_t h = (uint32_t)mix64((mix64(a[k].x) + mix64(a[k].y)) ^ hash);
		z[i].x = u[i] ^ h;
		z[i].y = (uint64_t)k << 32 | (uint32_t)(int32_t)u[i];
		k += (int32_t)u[i];
	}
	radix_sort_128x(z, z + n_u);
	for (i = 0; i < n_u >> 1; ++i) { mm128_t t = z[i]; z[i] = z[n_u-1-i], z[n_u-1-i] = t; }
	r = (mm_reg1_t*)calloc(n_u, sizeof(mm_reg1_t));
	for (i = 0; i < n_u; ++i) {
		mm_reg1_t *ri = &r[i];
		ri->id = i, ri->parent = MM_PARENT_UNSET;
		ri->score = ri->score0 = (int32_t)(z[i].x >> 32);
		ri->hash = (uint32_t)z[i].x;
		ri->cnt = (int32_t)z[i].y, ri->as = (int32_t)(z[i].y >> 32);
		ri->div = -1.0f;
		reg_set_coor(ri, qlen, a);
	}
	mm_afree(z);
	return r;
}

MM_FN void mm_split_reg(mm_reg1_t *r, mm_reg1_t *r2, int n, int qlen, mm128_t *a)
{ /* hit.c:90-107 */
	if (n <= 0 || n >= r->cnt) return;
	*r2 = *r;
	r2->id = -1, r2->sam_pri = 0, r2->p = 0, r2->split_inv = 0;
	r2->cnt = r->cnt - n;
	r2->score = (int32_t)(r->score * ((float)r2->cnt / r->cnt) + .499);
	r2->as = r->as + n;
	if (r->parent == r->id) r2->parent = MM_PARENT_TMP_PRI;
	reg_set_coor(r2, qlen, a);
	r->cnt -= r2->cnt, r->score -= r2->score;
	reg_set_coor(r, qlen, a);
	r->split |= 1, r2->split |= 2;
}

MM_FN void mm_set_parent(float mask_level, int n, mm_reg1_t *r, int sub_diff, int hard_mask_level)
{ /* hit.c:109-167 */
	int i, j, k, *w;
	uint64_t *cov;
	if (n <= 0) return;
	for (i = 0; i < n; ++i) r[i].id = i;
	cov = (uint64_t*)mm_amalloc((size_t)n * sizeof(uint64_t));
	w = (int*)mm_amalloc((size_t)n * sizeof(int));
	w[0] = 0, r[0].parent = 0;
	for (i = 1, k = 1; i < n; ++i) {
		mm_reg1_t *ri = &r[i];
		const int si = ri->qs, ei = ri->qe;
		int n_cov = 0, uncov_len = 0;
		if (!hard_mask_level) {
			for (j = 0; j < k; ++j) { /* query intervals of the primaries found so far that overlap hit i */
				const mm_reg1_t *rp = &r[w[j]];
				int sj = rp->qs, ej = rp->qe;
				if (ej <= si || sj >= ei) continue;
				if (sj < si) sj = si;
				if (ej > ei) ej = ei;
				cov[n_cov++] = (uint64_t)sj << 32 | (uint32_t)ej;
			}
			if (n_cov == 0) goto new_primary; /* j == k here */
			{ /* length of hit i not covered by any primary */
				int x = si;
				radix_sort_64(cov, cov + n_cov);
				for (j = 0; j < n_cov; ++j) {
					if ((int)(cov[j] >> 32) > x) uncov_len += (int)(cov[j] >> 32) - x;
					x = (int32_t)cov[j] > x ? (int32_t)cov[j] : x;
				}
				if (ei > x) uncov_len += ei - x;
			}
		}
		for (j = 0; j < k; ++j) {
			mm_reg1_t *rp = &r[w[j]];
			const int sj = rp->qs, ej = rp->qe;
			int min, max, ol;
			if (ej <= si || sj >= ei) continue;
			min = ej - sj < ei - si ? ej - sj : ei - si;
			max = ej - sj > ei - si ? ej - sj : ei - si;
			ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj) : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
			if ((float)ol / min - (float)uncov_len / max > mask_level) {
				int cnt_sub = 0;
				ri->parent = rp->parent;
				rp->subsc = rp->subsc > ri->score ? rp->subsc : ri->score;
				if (ri->cnt >= rp->cnt) cnt_sub = 1;
				if (rp->p && ri->p && (rp->rid != ri->rid || rp->rs != ri->rs || rp->re != ri->re || ol != min)) {
					rp->p->dp_max2 = rp->p->dp_max2 > ri->p->dp_max ? rp->p->dp_max2 : ri->p->dp_max;
					if (rp->p->dp_max - ri->p->dp_max <= sub_diff) cnt_sub = 1;
				}
				if (cnt_sub) ++rp->n_sub;
				break;
			}
		}
new_primary:
		if (j == k) w[k++] = i, ri->parent = i, ri->n_sub = 0;
	}
	mm_afree(cov); mm_afree(w);
}

MM_FN void mm_hit_sort(int *n_regs, mm_reg1_t *r)
{ /* hit.c:169-201: by dp_max (or chain score) then hash, descending; drops soft-deleted hits */
	int32_t i, n_aux, n = *n_regs, has_cigar = 0, no_cigar = 0;
	mm128_t *aux;
	mm_reg1_t *t;
	if (n <= 1) return;
	aux = (mm128_t*)mm_amalloc((size_t)n * 16);
	t = (mm_reg1_t*)mm_amalloc((size_t)n * sizeof(mm_reg1_t));
	for (i = n_aux = 0; i < n; ++i) {
		if (r[i].inv || r[i].cnt > 0) {
			if (r[i].p) aux[n_aux].x = (uint64_t)r[i].p->dp_max << 32 | r[i].hash, has_cigar = 1;
			else aux[n_aux].x = (uint64_t)r[i].score << 32 | r[i].hash, no_cigar = 1;
			aux[n_aux++].y = i;
		} else if (r[i].p) { free(r[i].p); r[i].p = 0; }
	}
	assert(has_cigar + no_cigar == 1);
	radix_sort_128x(aux, aux + n_aux);
	for (i = n_aux - 1; i >= 0; --i) t[n_aux - 1 - i] = r[aux[i].y];
	memcpy(r, t, sizeof(mm_reg1_t) * n_aux);
	*n_regs = n_aux;
	mm_afree(aux); mm_afree(t);
}

MM_FN int mm_set_sam_pri(int n, mm_reg1_t *r)
{ /* hit.c:203-212 */
	int i, n_pri = 0;
	for (i = 0; i < n; ++i) {
		if (r[i].id == r[i].parent) r[i].sam_pri = (++n_pri == 1);
		else r[i].sam_pri = 0;
	}
	return n_pri;
}

MM_FN void mm_sync_regs(int n_regs, mm_reg1_t *regs)
{ /* hit.c:214-236: renumber ids after removals and remap parents */
	int *map, i, max_id = -1, n_map;
	if (n_regs <= 0) return;
	for (i = 0; i < n_regs; ++i) max_id = max_id > regs[i].id ? max_id : regs[i].id;
	n_map = max_id + 1;
	map = (int*)mm_amalloc((size_t)(n_map > 0 ? n_map : 1) * sizeof(int));
	for (i = 0; i < n_map; ++i) map[i] = -1;
	for (i = 0; i < n_regs; ++i) if (regs[i].id >= 0) map[regs[i].id] = i;
	for (i = 0; i < n_regs; ++i) {
		mm_reg1_t *r = &regs[i];
		r->id = i;
		if (r->parent == MM_PARENT_TMP_PRI) r->parent = i;
		else if (r->parent >= 0 && map[r->parent] >= 0) r->parent = map[r->parent];
		else r->parent = MM_PARENT_UNSET;
	}
	mm_afree(map);
	mm_set_sam_pri(n_regs, regs);
}

MM_FN void mm_select_sub(float pri_ratio, int min_diff, int best_n, int *n_, mm_reg1_t *r)
{ /* hit.c:238-255 */
	int i, k, n = *n_, n_2nd = 0;
	if (!(pri_ratio > 0.0f && n > 0)) return;
	for (i = k = 0; i < n; ++i) {
		const int p = r[i].parent;
		if (p == i || r[i].inv) { r[k++] = r[i]; continue; }
		if ((r[i].score >= r[p].score * pri_ratio || r[i].score + min_diff >= r[p].score) && n_2nd < best_n) {
			if (!(r[i].qs == r[p].qs && r[i].qe == r[p].qe && r[i].rid == r[p].rid && r[i].rs == r[p].rs && r[i].re == r[p].re)) {
				r[k++] = r[i], ++n_2nd;
				continue;
			}
		}
		if (r[i].p) free(r[i].p);
	}
	if (k != n) mm_sync_regs(k, r);
	*n_ = k;
}

MM_FN void mm_select_sub_multi(float pri_ratio, float pri1, float pri2, int max_gap_ref, int min_diff, int best_n, int n_segs, const int *qlens, int *n_, mm_reg1_t *r)
{ /* pe.c:6-43 */
	int i, k, n = *n_, n_2nd = 0;
	const int max_dist = n_segs == 2 ? qlens[0] + qlens[1] + max_gap_ref : 0;
	if (!(pri_ratio > 0.0f && n > 0)) return;
	for (i = k = 0; i < n; ++i) {
		int keep = 0;
		if (r[i].parent == i) keep = 1;
		else if (r[i].score + min_diff >= r[r[i].parent].score) keep = 1;
		else {
			const mm_reg1_t *p = &r[r[i].parent], *q = &r[i];
			if (p->rev == q->rev && p->rid == q->rid && q->re - p->rs < max_dist && p->re - q->rs < max_dist) {
				if (q->score >= p->score * pri1) keep = 1; /* child near its parent on the reference */
			} else {
				const int par_both = (n_segs == 2 && p->qs < qlens[0] && p->qe > qlens[0]);
				const int chi_both = (n_segs == 2 && q->qs < qlens[0] && q->qe > qlens[0]);
				if (chi_both || chi_both == par_both) { if (q->score >= p->score * pri_ratio) keep = 1; }
				else if (q->score >= p->score * pri2) keep = 1;
			}
		}
		if (keep && r[i].parent != i && n_2nd++ >= best_n) keep = 0;
		if (keep) r[k++] = r[i];
		else if (r[i].p) free(r[i].p);
	}
	if (k != n) mm_sync_regs(k, r);
	*n_ = k;
}

MM_FN void mm_filter_regs(const mm_mapopt_t *opt, int qlen, int *n_regs, mm_reg1_t *regs)
{ /* hit.c:257-276 */
	int i, k;
	for (i = k = 0; i < *n_regs; ++i) {
		mm_reg1_t *r = &regs[i];
		int flt = 0;
		if (!r->inv && !r->seg_split && r->cnt < opt->min_cnt) flt = 1;
		if (r->p) {
			if (r->mlen < opt->min_chain_score) flt = 1;
			else if (r->p->dp_max < opt->min_dp_max) flt = 1;
			else if (r->qs > qlen * opt->max_clip_ratio && qlen - r->qe > qlen * opt->max_clip_ratio) flt = 1;
			if (flt) free(r->p);
		}
		if (flt) continue;
		if (k < i) regs[k] = regs[i];
		++k;
	}
	*n_regs = k;
}

MM_FN int mm_squeeze_a(int n_regs, mm_reg1_t *regs, mm128_t *a)
{ /* hit.c:278-296: compact a[] to the anchors still referenced, in order of `as` */
	int i, as = 0;
	uint64_t *aux = (uint64_t*)mm_amalloc((size_t)(n_regs > 0 ? n_regs : 1) * 8);
	for (i = 0; i < n_regs; ++i) aux[i] = (uint64_t)regs[i].as << 32 | (uint32_t)i;
	radix_sort_64(aux, aux + n_regs);
	for (i = 0; i < n_regs; ++i) {
		mm_reg1_t *r = &regs[(int32_t)aux[i]];
		if (r->as != as) {
			memmove(&a[as], &a[r->as], (size_t)r->cnt * 16);
			r->as = as;
		}
		as += r->cnt;
	}
	mm_afree(aux);
	return as;
}

MM_FN void mm_join_long(const mm_mapopt_t *opt, int qlen, int *n_regs_, mm_reg1_t *regs, mm128_t *a)
{ /* hit.c:298-354 */
	int i, n_aux, n_regs = *n_regs_, n_drop = 0;
	uint64_t *aux;
	if (n_regs < 2) return;
	mm_squeeze_a(n_regs, regs, a);
	aux = (uint64_t*)mm_amalloc((size_t)n_regs * 8);
	for (i = n_aux = 0; i < n_regs; ++i)
		if (regs[i].parent == i || regs[i].parent < 0) aux[n_aux++] = (uint64_t)regs[i].as << 32 | (uint32_t)i;
	radix_sort_64(aux, aux + n_aux);
	for (i = n_aux - 1; i >= 1; --i) {
		mm_reg1_t *r0 = &regs[(int32_t)aux[i-1]], *r1 = &regs[(int32_t)aux[i]];
		const mm128_t *a0e, *a1s;
		int max_gap, min_gap, sc_thres, min_flank_len;
		if (r0->as + r0->cnt != r1->as) continue;
		if (r0->rid != r1->rid || r0->rev != r1->rev) continue;
		a0e = &a[r0->as + r0->cnt - 1], a1s = &a[r1->as];
		if (a1s->x <= a0e->x || QPOS(*a1s) <= QPOS(*a0e)) continue;
		max_gap = min_gap = QPOS(*a1s) - QPOS(*a0e);
		max_gap = a0e->x + max_gap > a1s->x ? max_gap : (int)(a1s->x - a0e->x);
		min_gap = a0e->x + min_gap < a1s->x ? min_gap : (int)(a1s->x - a0e->x);
		if (max_gap > opt->max_join_long || min_gap > opt->max_join_short) continue;
		sc_thres = (int)((float)opt->min_join_flank_sc / opt->max_join_long * max_gap + .499);
		if (r0->score < sc_thres || r1->score < sc_thres) continue;
		min_flank_len = (int)(max_gap * opt->min_join_flank_ratio);
		if (r0->re - r0->rs < min_flank_len || r0->qe - r0->qs < min_flank_len) continue;
		if (r1->re - r1->rs < min_flank_len || r1->qe - r1->qs < min_flank_len) continue;
		a[r1->as].y |= MM_SEED_LONG_JOIN;
		r0->cnt += r1->cnt, r0->score += r1->score;
		reg_set_coor(r0, qlen, a);
		r1->cnt = 0;
		r1->parent = r0->id;
		++n_drop;
	}
	mm_afree(aux);
	if (n_drop > 0) {
		for (i = 0; i < n_regs; ++i) { /* re-point secondaries of a dropped chain to the surviving one */
			mm_reg1_t *r = &regs[i];
			if (r->parent >= 0 && r->id != r->parent)
				if (regs[r->parent].parent >= 0 && regs[r->parent].parent != r->parent)
					r->parent = regs[r->parent].parent;
		}
		mm_filter_regs(opt, qlen, n_regs_, regs);
		mm_sync_regs(*n_regs_, regs);
	}
}

MM_FN mm_seg_t *mm_seg_gen(uint32_t hash, int n_segs, const int *qlens, int n_regs0, const mm_reg1_t *regs0, int *n_regs, mm_reg1_t **regs, const mm128_t *a)
{ /* hit.c:356-410: split fragment chains into one chain list per segment, re-basing query coordinates */
	int s, i, j, acc_qlen[MM_MAX_SEG+1], qlen_sum;
	mm_seg_t *seg;
	assert(n_segs <= MM_MAX_SEG);
	for (s = 1, acc_qlen[0] = 0; s < n_segs; ++s) acc_qlen[s] = acc_qlen[s-1] + qlens[s-1];
	qlen_sum = acc_qlen[n_segs - 1] + qlens[n_segs - 1];
	seg = (mm_seg_t*)mm_acalloc(n_segs, sizeof(mm_seg_t));
	for (s = 0; s < n_segs; ++s) {
		seg[s].u = (uint64_t*)mm_amalloc((size_t)(n_regs0 > 0 ? n_regs0 : 1) * 8);
		for (i = 0; i < n_regs0; ++i) seg[s].u[i] = (uint64_t)regs0[i].score << 32;
	}
	for (i = 0; i < n_regs0; ++i)
		for (j = 0; j < regs0[i].cnt; ++j) {
			const int sid = (int)((a[regs0[i].as + j].y & MM_SEED_SEG_MASK) >> MM_SEED_SEG_SHIFT);
			++seg[sid].u[i], ++seg[sid].n_a;
		}
	for (s = 0; s < n_segs; ++s) {
		mm_seg_t *sr = &seg[s];
		for (i = 0, sr->n_u = 0; i < n_regs0; ++i)
			if ((int32_t)sr->u[i] != 0) sr->u[sr->n_u++] = sr->u[i];
		sr->a = (mm128_t*)mm_amalloc((size_t)(sr->n_a > 0 ? sr->n_a : 1) * sizeof(mm128_t));
		sr->n_a = 0;
	}
	for (i = 0; i < n_regs0; ++i)
		for (j = 0; j < regs0[i].cnt; ++j) {
			mm128_t a1 = a[regs0[i].as + j];
			const int sid = (int)((a1.y & MM_SEED_SEG_MASK) >> MM_SEED_SEG_SHIFT);
			a1.y -= a1.x >> 63 ? qlen_sum - (qlens[sid] + acc_qlen[sid]) : acc_qlen[sid];
			seg[sid].a[seg[sid].n_a++] = a1;
		}
	for (s = 0; s < n_segs; ++s) {
		regs[s] = mm_gen_regs(hash, qlens[s], seg[s].n_u, seg[s].u, seg[s].a);
		n_regs[s] = seg[s].n_u;
		for (i = 0; i < n_regs[s]; ++i) regs[s][i].seg_split = 1, regs[s][i].seg_id = s;
	}
	return seg;
}

MM_FN void mm_seg_free(int n_segs, mm_seg_t *segs)
{
	int i;
	for (i = 0; i < n_segs; ++i) { mm_afree(segs[i].u); mm_afree(segs[i].a); }
	mm_afree(segs);
}

#ifndef MM_DEVICE_BUILD /* MAPQ, pairing and the divergence estimate call logf(): host only */
MM_FN static void set_inv_mapq(int n_regs, mm_reg1_t *regs)
{ /* hit.c:420-444: an inversion hit takes the smaller MAPQ of its flanking primaries */
	int i, n_aux;
	mm128_t *aux;
	if (n_regs < 3) return;
	for (i = 0; i < n_regs; ++i) if (regs[i].inv) break;
	if (i == n_regs) return;
	aux = (mm128_t*)mm_amalloc((size_t)n_regs * 16);
	for (i = n_aux = 0; i < n_regs; ++i)
		if (regs[i].parent == i || regs[i].parent < 0)
			aux[n_aux].y = i, aux[n_aux++].x = (uint64_t)regs[i].rid << 32 | (uint32_t)regs[i].rs;
	radix_sort_128x(aux, aux + n_aux);
	for (i = 1; i < n_aux - 1; ++i) {
		mm_reg1_t *inv = &regs[aux[i].y];
		if (inv->inv) {
			const mm_reg1_t *l = &regs[aux[i-1].y], *r = &regs[aux[i+1].y];
			inv->mapq = l->mapq < r->mapq ? l->mapq : r->mapq;
		}
	}
	mm_afree(aux);
}

MM_FN void mm_set_mapq(int n_regs, mm_reg1_t *regs, int min_chain_sc, int match_sc, int rep_len, int is_sr)
{ /* hit.c:446-491.  Expression shapes are kept: every product below is evaluated in float, left to right. */
	static const float q_coef = 40.0f;
	int64_t sum_sc = 0;
	float uniq_ratio;
	int i;
	if (n_regs == 0) return;
	for (i = 0; i < n_regs; ++i)
		if (regs[i].parent == regs[i].id) sum_sc += regs[i].score;
	uniq_ratio = (float)sum_sc / (sum_sc + rep_len);
	for (i = 0; i < n_regs; ++i) {
		mm_reg1_t *r = &regs[i];
		if (r->inv || r->parent != r->id) { r->mapq = 0; continue; }
		{
			int mapq, subsc;
			float pen_s1 = (r->score > 100 ? 1.0f : 0.01f * r->score) * uniq_ratio;
			float pen_cm = r->cnt > 10 ? 1.0f : 0.1f * r->cnt;
			pen_cm = pen_s1 < pen_cm ? pen_s1 : pen_cm;
			subsc = r->subsc > min_chain_sc ? r->subsc : min_chain_sc;
			if (r->p && r->p->dp_max2 > 0 && r->p->dp_max > 0) {
				float identity = (float)r->mlen / r->blen;
				float x = (float)r->p->dp_max2 * subsc / r->p->dp_max / r->score0;
				mapq = (int)(identity * pen_cm * q_coef * (1.0f - x * x) * logf((float)r->p->dp_max / match_sc));
				if (!is_sr) {
					int mapq_alt = (int)(6.02f * identity * identity * (r->p->dp_max - r->p->dp_max2) / match_sc + .499f);
					mapq = mapq < mapq_alt ? mapq : mapq_alt;
				}
			} else {
				float x = (float)subsc / r->score0;
				if (r->p) {
					float identity = (float)r->mlen / r->blen;
					mapq = (int)(identity * pen_cm * q_coef * (1.0f - x) * logf((float)r->p->dp_max / match_sc));
				} else mapq = (int)(pen_cm * q_coef * (1.0f - x) * logf(r->score));
			}
			mapq -= (int)(4.343f * logf(r->n_sub + 1) + .499f);
			mapq = mapq > 0 ? mapq : 0;
			r->mapq = mapq < 60 ? mapq : 60;
			if (r->p && r->p->dp_max > r->p->dp_max2 && r->mapq == 0) r->mapq = 1;
		}
	}
	set_inv_mapq(n_regs, regs);
}

/* ---- pairing (pe.c:45-177) */

MM_FN static void set_pe_thru(const int *qlens, int *n_regs, mm_reg1_t **regs)
{ /* both mates cover the same short fragment end to end */
	int s, i, n_pri[2] = {0, 0}, pri[2] = {-1, -1};
	for (s = 0; s < 2; ++s)
		for (i = 0; i < n_regs[s]; ++i)
			if (regs[s][i].id == regs[s][i].parent) ++n_pri[s], pri[s] = i;
	if (n_pri[0] == 1 && n_pri[1] == 1) {
		mm_reg1_t *p = &regs[0][pri[0]], *q = &regs[1][pri[1]];
		if (p->rid == q->rid && p->rev == q->rev && abs(p->rs - q->rs) < 3 && abs(p->re - q->re) < 3
			&& ((p->qs == 0 && qlens[1] - q->qe == 0) || (q->qs == 0 && qlens[0] - p->qe == 0)))
			p->pe_thru = q->pe_thru = 1;
	}
}

typedef struct { int s, rev; uint64_t key; mm_reg1_t *r; } pair_elem_t;
#define KEY_PAIR(v) ((v).key)

/* the reference sorts pair_arr_t with its radix sort (pe.c:67-68); same permutation here */
MM_FN static void pair_ins(pair_elem_t *b, pair_elem_t *e)
{
	pair_elem_t *i, *j;
	for (i = b + 1; i < e; ++i) {
		if (!(i->key < (i - 1)->key)) continue;
		pair_elem_t t = *i;
		for (j = i; j > b && t.key < (j - 1)->key; --j) *j = *(j - 1);
		*j = t;
	}
}
MM_FN static void pair_lvl(pair_elem_t *b, pair_elem_t *e, int sh)
{
	size_t cnt[256]; pair_elem_t *hd[256], *tl[256], *i; int d;
	memset(cnt, 0, sizeof(cnt));
	for (i = b; i != e; ++i) ++cnt[(i->key >> sh) & 0xff];
	for (d = 0, i = b; d < 256; ++d) hd[d] = i, i += cnt[d], tl[d] = i;
	for (d = 0; d < 256;) {
		int to;
		if (hd[d] == tl[d]) { ++d; continue; }
		to = (int)((hd[d]->key >> sh) & 0xff);
		if (to == d) { ++hd[d]; continue; }
		{ pair_elem_t carry = *hd[d], sw;
		  do { sw = carry; carry = *hd[to]; *hd[to]++ = sw; to = (int)((carry.key >> sh) & 0xff); } while (to != d);
		  *hd[d]++ = carry; }
	}
	if (sh == 0) return;
	sh = sh > 8 ? sh - 8 : 0;
	for (d = 0, i = b; d < 256; i = tl[d], ++d) {
		if (tl[d] - i > 64) pair_lvl(i, tl[d], sh);
		else if (tl[d] - i > 1) pair_ins(i, tl[d]);
	}
}
static void sort_pairs(pair_elem_t *b, pair_elem_t *e) { if (e - b <= 64) pair_ins(b, e); else pair_lvl(b, e, 56); }

MM_FN void mm_pair(int max_gap_ref, int pe_bonus, int sub_diff, int match_sc, const int *qlens, int *n_regs, mm_reg1_t **regs)
{ /* pe.c:76-177 */
	int i, j, s, n, last[2], dp_thres, segs = 0, max_idx[2];
	int64_t max;
	pair_elem_t *a;
	uint64_t *sc = 0; size_t n_sc = 0, m_sc = 0;
	a = (pair_elem_t*)mm_amalloc((size_t)(n_regs[0] + n_regs[1] + 1) * sizeof(pair_elem_t));
	for (s = n = 0, dp_thres = 0; s < 2; ++s) {
		int best = 0;
		for (i = 0; i < n_regs[s]; ++i) {
			a[n].s = s, a[n].r = &regs[s][i], a[n].rev = a[n].r->rev;
			a[n].key = (uint64_t)a[n].r->rid << 32 | a[n].r->rs << 1 | (s ^ a[n].rev);
			best = best > a[n].r->p->dp_max ? best : a[n].r->p->dp_max;
			++n, segs |= 1 << s;
		}
		dp_thres += best;
	}
	if (segs != 3) { mm_afree(a); return; } /* only one mate mapped */
	dp_thres -= pe_bonus;
	if (dp_thres < 0) dp_thres = 0;
	sort_pairs(a, a + n);
	max = -1, max_idx[0] = max_idx[1] = -1, last[0] = last[1] = -1;
	m_sc = (size_t)n; sc = (uint64_t*)mm_amalloc((m_sc ? m_sc : 1) * 8); /* kv_resize(n) then kv_push growth */
	for (i = 0; i < n; ++i) {
		if (a[i].key & 1) { /* a mate that closes a pair: scan back over candidate openers on the same strand */
			mm_reg1_t *q, *r;
			if (last[a[i].rev] < 0) continue;
			r = a[i].r, q = a[last[a[i].rev]].r;
			if (r->rid != q->rid || r->rs - q->re > max_gap_ref) continue;
			for (j = last[a[i].rev]; j >= 0; --j) {
				int64_t score;
				if (a[j].rev != a[i].rev || a[j].s == a[i].s) continue;
				q = a[j].r;
				if (r->rid != q->rid || r->rs - q->re > max_gap_ref) break;
				if (r->p->dp_max + q->p->dp_max < dp_thres) continue;
				score = (int64_t)(r->p->dp_max + q->p->dp_max) << 32 | (r->hash + q->hash);
				if (score > max) max = score, max_idx[a[j].s] = j, max_idx[a[i].s] = i;
				if (n_sc == m_sc) { const size_t old = m_sc * 8; m_sc = m_sc ? m_sc << 1 : 2; sc = (uint64_t*)mm_arealloc(sc, old, m_sc * 8); }
				sc[n_sc++] = (uint64_t)score;
			}
		} else last[a[i].rev] = i;
	}
	if (n_sc > 1) radix_sort_64(sc, sc + n_sc);
	if (n_sc > 0 && max > 0) {
		int n_sub = 0, mapq_pe;
		mm_reg1_t *r[2];
		r[0] = a[max_idx[0]].r, r[1] = a[max_idx[1]].r;
		r[0]->proper_frag = r[1]->proper_frag = 1;
		for (s = 0; s < 2; ++s) {
			if (r[s]->id != r[s]->parent) { /* promote the paired hit to primary */
				mm_reg1_t *p = &regs[s][r[s]->parent];
				for (i = 0; i < n_regs[s]; ++i)
					if (regs[s][i].parent == p->id) regs[s][i].parent = r[s]->id;
				p->mapq = 0;
			}
			if (!r[s]->sam_pri) {
				for (i = 0; i < n_regs[s]; ++i) regs[s][i].sam_pri = 0;
				r[s]->sam_pri = 1;
			}
		}
		mapq_pe = r[0]->mapq > r[1]->mapq ? r[0]->mapq : r[1]->mapq;
		for (i = 0; i < (int)n_sc; ++i)
			if ((sc[i] >> 32) + sub_diff >= (uint64_t)max >> 32) ++n_sub;
		if (n_sc > 1) {
			int mapq_pe_alt = (int)(6.02f * ((max >> 32) - (sc[n_sc - 2] >> 32)) / match_sc - 4.343f * logf(n_sub));
			mapq_pe = mapq_pe < mapq_pe_alt ? mapq_pe : mapq_pe_alt;
		}
		if (r[0]->mapq < mapq_pe) r[0]->mapq = (int)(.2f * r[0]->mapq + .8f * mapq_pe + .499f);
		if (r[1]->mapq < mapq_pe) r[1]->mapq = (int)(.2f * r[1]->mapq + .8f * mapq_pe + .499f);
		if (n_sc == 1) {
			if (r[0]->mapq < 2) r[0]->mapq = 2;
			if (r[1]->mapq < 2) r[1]->mapq = 2;
		} else if ((uint64_t)max >> 32 > sc[n_sc - 2] >> 32) {
			if (r[0]->mapq < 1) r[0]->mapq = 1;
			if (r[1]->mapq < 1) r[1]->mapq = 1;
		}
	}
	mm_afree(a); mm_afree(sc);
	set_pe_thru(qlens, n_regs, regs);
}

/* ---- divergence estimate from the minimizers a chain did and did not use (esterr.c:6-64) */

MM_FN static inline int32_t fwd_qpos(int32_t qlen, const mm128_t *a)
{
	int32_t x = QPOS(*a);
	if (a->x >> 63) x = qlen - 1 - (x + 1 - SPAN(*a));
	return x;
}

MM_FN void mm_est_err(const mm_idx_t *mi, int qlen, int n_regs, mm_reg1_t *regs, const mm128_t *a, int32_t n, const uint64_t *mini_pos)
{
	int i;
	uint64_t sum_k = 0;
	float avg_k;
	if (n == 0) return;
	for (i = 0; i < n; ++i) sum_k += mini_pos[i] >> 32 & 0xff;
	avg_k = (float)sum_k / n;
	for (i = 0; i < n_regs; ++i) {
		mm_reg1_t *r = &regs[i];
		int32_t st, en, j, k, n_match, n_tot, l_ref, L = 0, R = n - 1, x0;
		r->div = -1.0f;
		if (r->cnt == 0) continue;
		x0 = fwd_qpos(qlen, r->rev ? &a[r->as + r->cnt - 1] : &a[r->as]);
		st = -1;
		while (L <= R) { /* locate the chain's first minimizer among the query minimizers */
			const int32_t m = (int32_t)(((uint64_t)L + R) >> 1), y = (int32_t)mini_pos[m];
			if (y < x0) L = m + 1;
			else if (y > x0) R = m - 1;
			else { st = m; break; }
		}
		en = st;
		if (st < 0) {
			if (mm_verbose >= 2) fprintf(stderr, "[WARNING] logic inconsistency in mm_est_err(). Please contact the developer.\n");
			continue;
		}
		l_ref = mi->seq[r->rid].len;
		for (k = 1, j = st + 1, n_match = 1; j < n && k < r->cnt; ++j) {
			const int32_t x = fwd_qpos(qlen, r->rev ? &a[r->as + r->cnt - 1 - k] : &a[r->as + k]);
			if (x == (int32_t)mini_pos[j]) ++k, en = j, ++n_match;
		}
		n_tot = en - st + 1;
		if (r->qs > avg_k && r->rs > avg_k) ++n_tot;
		if (qlen - r->qs > avg_k && l_ref - r->re > avg_k) ++n_tot;
		r->div = logf((float)n_tot / n_match) / avg_k;
	}
}
#endif /* !MM_DEVICE_BUILD */
