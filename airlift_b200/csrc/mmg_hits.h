// mmg_hits.h -- chains -> hit records -> per-mate hit lists, MAPQ and pairing: the allocation-free building blocks of the
// post-chaining device stages (mmg_post.cu).  Everything works on caller-provided spans of plain records addressed by index;
// there is no heap, no pointer chasing and no per-fragment object.  __host__ __device__, so that tests/emu/ can hold the very
// code the kernels run against the reference's mm_map_frag on the CPU.
//
// What each block must reproduce (paths relative to src/minimap2-master_remapping/):
//   hit keys and order        mm_gen_regs            hit.c:52-88   (klib radix order, ties included)
//   coordinates               mm_reg_set_coor        hit.c:8-38
//   primary / secondary tree  mm_set_parent          hit.c:109-167
//   secondary selection       mm_select_sub(_multi)  hit.c:238-255, pe.c:6-43; renumbering mm_sync_regs hit.c:214-236
//   per-mate split            mm_seg_gen             hit.c:356-410
//   filters and final order   mm_filter_regs, mm_hit_sort   hit.c:257-276, 169-201
//   MAPQ                      mm_set_mapq            hit.c:446-491 (logf taken from a host-made table of glibc values)
//   pairing                   mm_pair, mm_set_pe_thru       pe.c:45-177
#pragma once
#include "mmg_core.h"
#include "mmg_warp.h"

// == mm_reg1_t (minimap.h:83-98), 80 bytes; `p` holds 1 + the index of the hit's alignment record instead of a pointer
struct HitRec {
	int32_t id, cnt, rid, score;
	int32_t qs, qe, rs, re;
	int32_t parent, subsc;
	int32_t as;
	int32_t mlen, blen;
	int32_t n_sub;
	int32_t score0;
	uint32_t bits;   // mapq:8 split:2 rev:1 inv:1 sam_pri:1 proper_frag:1 pe_thru:1 seg_split:1 seg_id:8 split_inv:1 (LSB first)
	uint32_t hash;
	float div;
	uint64_t p;
};
static_assert(sizeof(HitRec) == 80, "HitRec must have the layout of mm_reg1_t");

#define HB_MAPQ(b)      ((b) & 0xffu)
#define HB_SPLIT(b)     ((b) >> 8 & 3u)
#define HB_REV          (1u << 10)
#define HB_INV          (1u << 11)
#define HB_SAM_PRI      (1u << 12)
#define HB_PROPER       (1u << 13)
#define HB_PE_THRU      (1u << 14)
#define HB_SEG_SPLIT    (1u << 15)
#define HB_SEG_ID(b)    ((b) >> 16 & 0xffu)
#define HB_SPLIT_INV    (1u << 24)
#define HIT_PARENT_UNSET   (-1)
#define HIT_PARENT_TMP_PRI (-2)

// == mm_extra_t header (minimap.h:75-81), 24 bytes, followed by n_cigar words
struct HitExtra {
	uint32_t capacity;
	int32_t dp_score, dp_max, dp_max2;
	uint32_t n_ambi_strand;  // n_ambi:30, trans_strand:2
	uint32_t n_cigar;
};
static_assert(sizeof(HitExtra) == 24, "HitExtra must have the layout of mm_extra_t");

struct HitOpt {   // the fields of mm_mapopt_t / mm_idx_t these stages read
	int64_t flag;
	float mask_level, pri_ratio, max_clip_ratio;
	int32_t best_n, a, b, q, e, q2, e2, sc_ambi, zdrop, zdrop_inv, end_bonus, min_dp_max, min_cnt, min_chain_score, bw;
	int32_t pe_ori, pe_bonus, max_gap, max_gap_ref, max_frag_len, k, max_qlen;
	int64_t max_sw_mat;
};
#define HIT_F_CIGAR        0x004LL
#define HIT_F_ALL_CHAINS   0x800000LL
#define HIT_F_HARD_MLEVEL  0x20000000LL
#define HIT_F_SR           0x1000LL

MMG_HD HitExtra *hit_ext(uint32_t *xw, uint64_t p) { return reinterpret_cast<HitExtra*>(xw + (p - 1)); }

MMG_HD int32_t hit_span(const mm128 &an) { return (int32_t)(an.y >> 32 & 0xff); }
MMG_HD int32_t hit_rpos(const mm128 &an) { return (int32_t)an.x; }
MMG_HD int32_t hit_qpos(const mm128 &an) { return (int32_t)an.y; }

MMG_HD uint64_t hit_mix64(uint64_t key)  // hit.c:40-50
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

// sort key of a chain (hit.c:62-64): its score|count word salted with a hash of its first anchor
MMG_HD uint64_t hit_key(uint64_t u, const mm128 &first, uint32_t frag_hash)
{
	return u ^ (uint32_t)hit_mix64((hit_mix64(first.x) + hit_mix64(first.y)) ^ frag_hash);
}

// coordinates and fuzzy lengths of a hit from its chain (hit.c:8-38)
MMG_HD void hit_set_coor(HitRec *r, int32_t qlen, const mm128 *a)
{
	const mm128 first = a[r->as], last = a[r->as + r->cnt - 1];
	const int32_t q_span = hit_span(first);
	const bool rev = first.x >> 63;
	r->bits = rev ? (r->bits | HB_REV) : (r->bits & ~HB_REV);
	r->rid = (int32_t)(first.x << 1 >> 33);
	r->rs = hit_rpos(first) + 1 > q_span ? hit_rpos(first) + 1 - q_span : 0;
	r->re = hit_rpos(last) + 1;
	if (!rev) r->qs = hit_qpos(first) + 1 - q_span, r->qe = hit_qpos(last) + 1;
	else r->qs = qlen - (hit_qpos(last) + 1), r->qe = qlen - (hit_qpos(first) + 1 - q_span);
	r->mlen = r->blen = 0;
	if (r->cnt <= 0) return;
	int32_t ml = q_span, bl = q_span;
	mm128 prev = first;
	for (int32_t i = r->as + 1; i < r->as + r->cnt; ++i) {
		const mm128 cur = a[i];
		const int32_t span = hit_span(cur), tl = hit_rpos(cur) - hit_rpos(prev), ql = hit_qpos(cur) - hit_qpos(prev);
		bl += tl > ql ? tl : ql;
		ml += tl > span && ql > span ? span : tl < ql ? tl : ql;
		prev = cur;
	}
	r->mlen = ml, r->blen = bl;
}

// a fresh hit from its sorted key and (first anchor, count) word (hit.c:73-85)
MMG_HD void hit_init(HitRec *r, int32_t idx, uint64_t key, uint64_t as_cnt, int32_t qlen, const mm128 *a)
{
	r->id = idx, r->parent = HIT_PARENT_UNSET;
	r->score = r->score0 = (int32_t)(key >> 32);
	r->hash = (uint32_t)key;
	r->cnt = (int32_t)as_cnt, r->as = (int32_t)(as_cnt >> 32);
	r->subsc = 0, r->n_sub = 0, r->bits = 0, r->p = 0;
	r->div = -1.0f;
	hit_set_coor(r, qlen, a);
}

// ---- small in-place sorts (values only: any algorithm gives the reference's result)
MMG_HDN inline void hit_sort_u64(uint64_t *v, int n)
{ // heap sort, ascending
	for (int s = n / 2 - 1; s >= 0; --s) {
		int i = s; const uint64_t t = v[i];
		for (int k; (k = 2 * i + 1) < n; i = k) { if (k + 1 < n && v[k + 1] > v[k]) ++k; if (v[k] <= t) break; v[i] = v[k]; }
		v[i] = t;
	}
	for (int m = n - 1; m > 0; --m) {
		const uint64_t t = v[m]; v[m] = v[0];
		int i = 0;
		for (int k; (k = 2 * i + 1) < m; i = k) { if (k + 1 < m && v[k + 1] > v[k]) ++k; if (v[k] <= t) break; v[i] = v[k]; }
		v[i] = t;
	}
}

// ---- mm_set_parent (hit.c:109-167).  w[] (n ints) lists the primaries found so far; cov[] (n words) holds the clipped
// query intervals of the primaries that overlap the hit under test.  Serial in i by nature (a hit is judged against the
// primaries before it); the number of primaries of a read is small, so this is O(n) in practice.
// xw: the word array holding the alignment records (HitRec::p = 1 + word offset), or null before alignment.
MMG_HDN inline void hit_set_parent(float mask_level, int n, HitRec *r, int sub_diff, bool hard_mask_level, int32_t *w, uint64_t *cov, uint32_t *xw)
{
	if (n <= 0) return;
	for (int i = 0; i < n; ++i) r[i].id = i;
	w[0] = 0, r[0].parent = 0;
	int k = 1;
	for (int i = 1; i < n; ++i) {
		HitRec *ri = &r[i];
		const int si = ri->qs, ei = ri->qe;
		int n_cov = 0, uncov_len = 0, j = k;
		bool judge = true;
		if (!hard_mask_level) {
			for (int t = 0; t < k; ++t) {
				int sj = r[w[t]].qs, ej = r[w[t]].qe;
				if (ej <= si || sj >= ei) continue;
				if (sj < si) sj = si;
				if (ej > ei) ej = ei;
				cov[n_cov++] = (uint64_t)sj << 32 | (uint32_t)ej;
			}
			if (n_cov == 0) judge = false; // overlaps no primary: a new primary
			else {
				int x = si;
				if (n_cov > 1) hit_sort_u64(cov, n_cov);
				for (int t = 0; t < n_cov; ++t) {
					if ((int)(cov[t] >> 32) > x) uncov_len += (int)(cov[t] >> 32) - x;
					x = (int32_t)cov[t] > x ? (int32_t)cov[t] : x;
				}
				if (ei > x) uncov_len += ei - x;
			}
		}
		if (judge)
			for (j = 0; j < k; ++j) {
				HitRec *rp = &r[w[j]];
				const int sj = rp->qs, ej = rp->qe;
				if (ej <= si || sj >= ei) continue;
				const int mn = ej - sj < ei - si ? ej - sj : ei - si, mx = ej - sj > ei - si ? ej - sj : ei - si;
				const int ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj) : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
				if ((float)ol / mn - (float)uncov_len / mx > mask_level) {
					int cnt_sub = 0;
					ri->parent = rp->parent;
					rp->subsc = rp->subsc > ri->score ? rp->subsc : ri->score;
					if (ri->cnt >= rp->cnt) cnt_sub = 1;
					if (rp->p && ri->p && xw) {
						HitExtra *xp = hit_ext(xw, rp->p), *xi = hit_ext(xw, ri->p);
						if (rp->rid != ri->rid || rp->rs != ri->rs || rp->re != ri->re || ol != mn) {
							xp->dp_max2 = xp->dp_max2 > xi->dp_max ? xp->dp_max2 : xi->dp_max;
							if (xp->dp_max - xi->dp_max <= sub_diff) cnt_sub = 1;
						}
					}
					if (cnt_sub) ++rp->n_sub;
					break;
				}
			}
		if (j == k) w[k++] = i, ri->parent = i, ri->n_sub = 0;
	}
}

MMG_HD int hit_set_sam_pri(int n, HitRec *r)  // hit.c:203-212
{
	int n_pri = 0;
	for (int i = 0; i < n; ++i) {
		if (r[i].id == r[i].parent) { ++n_pri; r[i].bits = n_pri == 1 ? (r[i].bits | HB_SAM_PRI) : (r[i].bits & ~HB_SAM_PRI); }
		else r[i].bits &= ~HB_SAM_PRI;
	}
	return n_pri;
}

// renumber ids after removals and re-point parents (hit.c:214-236); map[] holds max_id + 1 ints
MMG_HDN inline void hit_sync(int n, HitRec *r, int32_t *map)
{
	if (n <= 0) return;
	int max_id = -1;
	for (int i = 0; i < n; ++i) max_id = max_id > r[i].id ? max_id : r[i].id;
	for (int i = 0; i <= max_id; ++i) map[i] = -1;
	for (int i = 0; i < n; ++i) if (r[i].id >= 0) map[r[i].id] = i;
	for (int i = 0; i < n; ++i) {
		HitRec *h = &r[i];
		h->id = i;
		if (h->parent == HIT_PARENT_TMP_PRI) h->parent = i;
		else if (h->parent >= 0 && map[h->parent] >= 0) h->parent = map[h->parent];
		else h->parent = HIT_PARENT_UNSET;
	}
	hit_set_sam_pri(n, r);
}

// hit.c:238-255.  Dropped hits simply leave the list (their alignment records are never referenced again).
MMG_HDN inline int hit_select_sub(float pri_ratio, int min_diff, int best_n, int n, HitRec *r, int32_t *map)
{
	if (!(pri_ratio > 0.0f && n > 0)) return n;
	int k = 0, n_2nd = 0;
	for (int i = 0; i < n; ++i) {
		const int p = r[i].parent;
		if (p == i || (r[i].bits & HB_INV)) { r[k++] = r[i]; continue; }
		if ((r[i].score >= r[p].score * pri_ratio || r[i].score + min_diff >= r[p].score) && n_2nd < best_n)
			if (!(r[i].qs == r[p].qs && r[i].qe == r[p].qe && r[i].rid == r[p].rid && r[i].rs == r[p].rs && r[i].re == r[p].re)) {
				r[k++] = r[i], ++n_2nd;
				continue;
			}
	}
	if (k != n) hit_sync(k, r, map);
	return k;
}

// pe.c:6-43
MMG_HDN inline int hit_select_sub_multi(float pri_ratio, float pri1, float pri2, int max_gap_ref, int min_diff, int best_n, int n_segs, const int32_t *qlens,
                                        int n, HitRec *r, int32_t *map)
{
	if (!(pri_ratio > 0.0f && n > 0)) return n;
	const int max_dist = n_segs == 2 ? qlens[0] + qlens[1] + max_gap_ref : 0;
	int k = 0, n_2nd = 0;
	for (int i = 0; i < n; ++i) {
		bool keep = false;
		if (r[i].parent == i) keep = true;
		else if (r[i].score + min_diff >= r[r[i].parent].score) keep = true;
		else {
			const HitRec *p = &r[r[i].parent], *q = &r[i];
			const bool prev = p->bits & HB_REV, qrev = q->bits & HB_REV;
			if (prev == qrev && p->rid == q->rid && q->re - p->rs < max_dist && p->re - q->rs < max_dist) {
				if (q->score >= p->score * pri1) keep = true;
			} else {
				const int par_both = (n_segs == 2 && p->qs < qlens[0] && p->qe > qlens[0]);
				const int chi_both = (n_segs == 2 && q->qs < qlens[0] && q->qe > qlens[0]);
				if (chi_both || chi_both == par_both) { if (q->score >= p->score * pri_ratio) keep = true; }
				else if (q->score >= p->score * pri2) keep = true;
			}
		}
		if (keep && r[i].parent != i && n_2nd++ >= best_n) keep = false;
		if (keep) r[k++] = r[i];
	}
	if (k != n) hit_sync(k, r, map);
	return k;
}

// hit.c:257-276; xw as in hit_set_parent
MMG_HDN inline int hit_filter(const HitOpt &o, int qlen, int n, HitRec *r, uint32_t *xw)
{
	int k = 0;
	for (int i = 0; i < n; ++i) {
		HitRec *h = &r[i];
		bool flt = false;
		if (!(h->bits & HB_INV) && !(h->bits & HB_SEG_SPLIT) && h->cnt < o.min_cnt) flt = true;
		if (h->p) {
			const HitExtra *x = hit_ext(xw, h->p);
			if (h->mlen < o.min_chain_score) flt = true;
			else if (x->dp_max < o.min_dp_max) flt = true;
			else if (h->qs > qlen * o.max_clip_ratio && qlen - h->qe > qlen * o.max_clip_ratio) flt = true;
		}
		if (flt) continue;
		if (k < i) r[k] = r[i];
		++k;
	}
	return k;
}

// The order klib's radix_sort_128x leaves n records in when their keys are sorted ascending and then read back to front
// (mm_gen_regs, mm_hit_sort): descending by key; up to 64 records the sort is a stable insertion sort (ksort.h:101-151 with
// RS_MIN_SIZE 64), so equal keys come out in reverse input order.  Larger inputs with equal keys replay the radix sort.
// key[] / idx[] (n each) are scratch; on return idx[rank] = input index.
MMG_HDN inline void hit_order_desc(int n, uint64_t *key, int32_t *idx, mm128 *big /* n, only touched when n > 64 */, RsFrame *stack)
{
	if (n <= 64) {
		for (int i = 0; i < n; ++i) idx[i] = i;
		for (int i = 1; i < n; ++i) { // insertion sort, descending by (key, input index)
			const uint64_t kk = key[idx[i]]; const int32_t ii = idx[i];
			int j = i;
			for (; j > 0 && (key[idx[j - 1]] < kk || (key[idx[j - 1]] == kk && idx[j - 1] < ii)); --j) idx[j] = idx[j - 1];
			idx[j] = ii;
		}
		return;
	}
	for (int i = 0; i < n; ++i) big[i].x = key[i], big[i].y = (uint64_t)i;
	mmg_rs_sort_exact(big, (int64_t)n, stack, KeyX());
	for (int i = 0; i < n; ++i) idx[i] = (int32_t)big[n - 1 - i].y;
}


// mm_hit_sort (hit.c:169-201): drop emptied hits, order by dp_max (or chain score) then hash, descending.
// key/idx: n scratch entries each; tmp: n records; big/stack as in hit_order_desc.
MMG_HDN inline int hit_final_sort(int n, HitRec *r, uint32_t *xw, uint64_t *key, int32_t *idx, HitRec *tmp, mm128 *big, RsFrame *stack)
{
	if (n <= 1) return n;
	int m = 0;
	for (int i = 0; i < n; ++i)
		if ((r[i].bits & HB_INV) || r[i].cnt > 0) {
			key[m] = r[i].p ? ((uint64_t)(uint32_t)hit_ext(xw, r[i].p)->dp_max << 32 | r[i].hash) : ((uint64_t)(uint32_t)r[i].score << 32 | r[i].hash);
			tmp[m++] = r[i];
		}
	hit_order_desc(m, key, idx, big, stack);
	for (int i = 0; i < m; ++i) r[i] = tmp[idx[i]];
	return m;
}

// ---- mm_seg_gen (hit.c:356-410): the hits of a fragment, split into one hit list per mate.
// Every anchor of a kept chain goes to the list of its own segment with its query coordinate re-based to that read;
// a fragment chain becomes one chain per segment it touches.  out_a: the fragment's share of the per-mate anchor arena
// (as many anchors as the kept chains hold), segment 0 first; seg_a0[s] = first anchor of segment s inside it.
// u_s / key / idx / big: scratch of n0 entries each.  Returns the per-mate hit counts in n_out[].
MMG_HDN inline void hit_split_mates(uint32_t frag_hash, int n_segs, const int32_t *qlens, int n0, const HitRec *r0, const mm128 *a, mm128 *out_a, int32_t *seg_a0,
                                    HitRec *const *r_out, int32_t *n_out, uint64_t *u_s, uint64_t *key, int32_t *idx, mm128 *big, RsFrame *stack)
{
	int32_t acc[9], n_a[8], qlen_sum;
	acc[0] = 0;
	for (int s = 1; s < n_segs; ++s) acc[s] = acc[s - 1] + qlens[s - 1];
	qlen_sum = acc[n_segs - 1] + qlens[n_segs - 1];
	for (int s = 0; s < n_segs; ++s) n_a[s] = 0;
	for (int i = 0; i < n0; ++i)
		for (int j = 0; j < r0[i].cnt; ++j) ++n_a[(a[r0[i].as + j].y & MMG_SEED_SEG_MASK) >> MMG_SEED_SEG_SHIFT];
	seg_a0[0] = 0;
	for (int s = 1; s <= n_segs; ++s) seg_a0[s] = seg_a0[s - 1] + n_a[s - 1];
	for (int s = 0; s < n_segs; ++s) {
		mm128 *as = out_a + seg_a0[s];
		int n_u = 0, fill = 0;
		for (int i = 0; i < n0; ++i) { // chains in hit order; anchors keep their order inside a chain
			int c = 0;
			for (int j = 0; j < r0[i].cnt; ++j) {
				mm128 a1 = a[r0[i].as + j];
				if ((int)((a1.y & MMG_SEED_SEG_MASK) >> MMG_SEED_SEG_SHIFT) != s) continue;
				a1.y -= (uint64_t)(a1.x >> 63 ? qlen_sum - (qlens[s] + acc[s]) : acc[s]);
				as[fill++] = a1, ++c;
			}
			if (c) u_s[n_u++] = (uint64_t)(uint32_t)r0[i].score << 32 | (uint32_t)c;
		}
		// mm_gen_regs on the mate's chains (hit.c:52-88)
		int k = 0;
		for (int i = 0; i < n_u; ++i) { key[i] = hit_key(u_s[i], as[k], frag_hash); k += (int32_t)u_s[i]; }
		hit_order_desc(n_u, key, idx, big, stack);
		// first anchor of chain i = prefix sum of the counts before it: recompute per output hit (n_u is small)
		for (int o = 0; o < n_u; ++o) {
			const int i = idx[o];
			int first = 0;
			for (int t = 0; t < i; ++t) first += (int32_t)u_s[t];
			HitRec *h = &r_out[s][o];
			hit_init(h, o, key[i], (uint64_t)first << 32 | (uint32_t)(int32_t)u_s[i], qlens[s], as);
			h->bits |= HB_SEG_SPLIT | ((uint32_t)s << 16);
		}
		n_out[s] = n_u;
	}
}

// mm_squeeze_a (hit.c:278-296) for a single-segment read: anchors of the kept hits move to the front of a[], in order of
// their first anchor.  key: n scratch words.  Returns the number of anchors kept.
MMG_HDN inline int hit_squeeze(int n, HitRec *r, mm128 *a, uint64_t *key)
{
	int as = 0;
	for (int i = 0; i < n; ++i) key[i] = (uint64_t)(uint32_t)r[i].as << 32 | (uint32_t)i;
	if (n > 1) hit_sort_u64(key, n);
	for (int i = 0; i < n; ++i) {
		HitRec *h = &r[(int32_t)(uint32_t)key[i]];
		if (h->as != as) {
			for (int j = 0; j < h->cnt; ++j) a[as + j] = a[h->as + j]; // as < h->as: a forward copy never overwrites unread anchors
			h->as = as;
		}
		as += h->cnt;
	}
	return as;
}

// ---- MAPQ (hit.c:446-491).  logf comes from two host-made tables of glibc values: ld[i] = logf((float)i / match_sc),
// li[i] = logf((float)i); every argument the reference passes is one of these.  Returns false if an argument
// lies outside the tables (the caller reports it; nothing is approximated).
struct LogTab { const float *ld, *li; int32_t n; };

MMG_HDN inline bool hit_set_mapq(int n, HitRec *r, uint32_t *xw, int min_chain_sc, int match_sc, int rep_len, bool is_sr, const LogTab &lt, uint64_t *key /* n scratch */)
{
	const float q_coef = 40.0f;
	if (n == 0) return true;
	int64_t sum_sc = 0;
	for (int i = 0; i < n; ++i) if (r[i].parent == r[i].id) sum_sc += r[i].score;
	const float uniq_ratio = (float)sum_sc / (sum_sc + rep_len);
	bool ok = true, any_inv = false;
	for (int i = 0; i < n; ++i) {
		HitRec *h = &r[i];
		if (h->bits & HB_INV) any_inv = true;
		if ((h->bits & HB_INV) || h->parent != h->id) { h->bits &= ~0xffu; continue; }
		const HitExtra *x = h->p ? hit_ext(xw, h->p) : nullptr;
		int mapq;
		const float pen_s1 = (h->score > 100 ? 1.0f : 0.01f * h->score) * uniq_ratio;
		float pen_cm = h->cnt > 10 ? 1.0f : 0.1f * h->cnt;
		pen_cm = pen_s1 < pen_cm ? pen_s1 : pen_cm;
		const int subsc = h->subsc > min_chain_sc ? h->subsc : min_chain_sc;
		if (x && x->dp_max2 > 0 && x->dp_max > 0) {
			const float identity = (float)h->mlen / h->blen;
			const float xx = (float)x->dp_max2 * subsc / x->dp_max / h->score0;
			if (x->dp_max >= lt.n) { ok = false; continue; }
			mapq = (int)(identity * pen_cm * q_coef * (1.0f - xx * xx) * lt.ld[x->dp_max]);
			if (!is_sr) {
				const int mapq_alt = (int)(6.02f * identity * identity * (x->dp_max - x->dp_max2) / match_sc + .499f);
				mapq = mapq < mapq_alt ? mapq : mapq_alt;
			}
		} else {
			const float xx = (float)subsc / h->score0;
			if (x) {
				const float identity = (float)h->mlen / h->blen;
				if (x->dp_max < 0 || x->dp_max >= lt.n) { ok = false; continue; }
				mapq = (int)(identity * pen_cm * q_coef * (1.0f - xx) * lt.ld[x->dp_max]);
			} else {
				if (h->score < 0 || h->score >= lt.n) { ok = false; continue; }
				mapq = (int)(pen_cm * q_coef * (1.0f - xx) * lt.li[h->score]);
			}
		}
		if (h->n_sub + 1 >= lt.n) { ok = false; continue; }
		mapq -= (int)(4.343f * lt.li[h->n_sub + 1] + .499f);
		mapq = mapq > 0 ? mapq : 0;
		mapq = mapq < 60 ? mapq : 60;
		if (x && x->dp_max > x->dp_max2 && mapq == 0) mapq = 1;
		h->bits = (h->bits & ~0xffu) | (uint32_t)mapq;
	}
	if (any_inv && n >= 3) { // mm_set_inv_mapq (hit.c:420-444): an inversion takes the smaller MAPQ of the primaries around it on the reference
		int m = 0;
		for (int i = 0; i < n; ++i)
			if (r[i].parent == i || r[i].parent < 0) key[m++] = (uint64_t)(uint32_t)r[i].rid << 52 | (uint64_t)(uint32_t)r[i].rs << 20 | (uint32_t)i; // rid < 4096, n < 2^20: holds for the ordering below
		bool fits = true;
		for (int i = 0; i < n; ++i) if (r[i].rid >= 4096 || n >= (1 << 20)) fits = false;
		if (!fits) return false;
		if (m > 1) hit_sort_u64(key, m);
		for (int t = 1; t < m - 1; ++t) {
			HitRec *inv = &r[key[t] & 0xfffff];
			if (inv->bits & HB_INV) {
				const uint32_t l = HB_MAPQ(r[key[t - 1] & 0xfffff].bits), rr = HB_MAPQ(r[key[t + 1] & 0xfffff].bits);
				inv->bits = (inv->bits & ~0xffu) | (l < rr ? l : rr);
			}
		}
	}
	return ok;
}

// ---- pairing (pe.c:45-177).  The hits of both mates are visited in the reference's order: ascending by (contig, start,
// closes-a-pair bit).  el[] holds that key per hit, el2[] = segment << 30 | strand << 29 | index, idx[] the sorted order.
struct PairScan {
	int n; const uint64_t *el; const int32_t *el2, *idx; HitRec *const *regs; uint32_t *xw; int max_gap_ref, dp_thres;
	MMG_HD int seg(int t) const { return (int)((uint32_t)el2[idx[t]] >> 30 & 1); }
	MMG_HD int rev(int t) const { return (int)((uint32_t)el2[idx[t]] >> 29 & 1); }
	MMG_HD HitRec *hit(int t) const { return &regs[seg(t)][el2[idx[t]] & 0x1fffffff]; }
	// every (opener j, closer i) the scan of pe.c:113-131 scores, in its order
	template <class F> MMG_HD void each(F &&f) const
	{
		int last[2] = {-1, -1};
		for (int i = 0; i < n; ++i) {
			if (!(el[idx[i]] & 1)) { last[rev(i)] = i; continue; }
			const int rv = rev(i);
			if (last[rv] < 0) continue;
			const HitRec *r = hit(i), *q = hit(last[rv]);
			if (r->rid != q->rid || r->rs - q->re > max_gap_ref) continue;
			for (int j = last[rv]; j >= 0; --j) {
				if (rev(j) != rv || seg(j) == seg(i)) continue;
				q = hit(j);
				if (r->rid != q->rid || r->rs - q->re > max_gap_ref) break;
				const int dsum = hit_ext(xw, r->p)->dp_max + hit_ext(xw, q->p)->dp_max;
				if (dsum < dp_thres) continue;
				f(i, j, (int64_t)dsum << 32 | (uint32_t)(r->hash + q->hash));
			}
		}
	}
};

MMG_HDN inline bool hit_pair(int max_gap_ref, int pe_bonus, int sub_diff, int match_sc, const int32_t *qlens, const int32_t *n_regs, HitRec *const *regs, uint32_t *xw,
                             const LogTab &lt, uint64_t *el, int32_t *el2, int32_t *idx, mm128 *big, RsFrame *stack)
{
	int n = 0, segs = 0, dp_thres = 0;
	bool ok = true;
	for (int s = 0; s < 2; ++s) {
		int best = 0;
		for (int i = 0; i < n_regs[s]; ++i) {
			const HitRec *h = &regs[s][i];
			const int rev = (h->bits & HB_REV) ? 1 : 0;
			const int32_t lo = (int32_t)((uint32_t)h->rs << 1) | (s ^ rev); // `r->rs << 1 | (s ^ rev)` is an int in the reference and gets sign-extended
			el[n] = (uint64_t)(uint32_t)h->rid << 32 | (uint64_t)(int64_t)lo;
			el2[n] = (int32_t)((uint32_t)s << 30 | (uint32_t)rev << 29 | (uint32_t)i);
			const int dm = hit_ext(xw, h->p)->dp_max;
			best = best > dm ? best : dm;
			++n, segs |= 1 << s;
		}
		dp_thres += best;
	}
	if (segs != 3) return true; // only one mate has hits: nothing to pair, and the reference leaves before mm_set_pe_thru
	{
		dp_thres -= pe_bonus;
		if (dp_thres < 0) dp_thres = 0;
		if (n <= 64) { // radix_sort_pair (pe.c:67-68): klib's order, a stable insertion sort up to 64 elements
			for (int i = 0; i < n; ++i) idx[i] = i;
			for (int i = 1; i < n; ++i) {
				const int32_t ii = idx[i]; const uint64_t kk = el[ii];
				int j = i;
				for (; j > 0 && kk < el[idx[j - 1]]; --j) idx[j] = idx[j - 1];
				idx[j] = ii;
			}
		} else {
			for (int i = 0; i < n; ++i) big[i].x = el[i], big[i].y = (uint64_t)i;
			mmg_rs_sort_exact(big, (int64_t)n, stack, KeyX());
			for (int i = 0; i < n; ++i) idx[i] = (int32_t)big[i].y;
		}
		PairScan ps = {n, el, el2, idx, regs, xw, max_gap_ref, dp_thres};
		// the reference collects every pair score, sorts them, and reads the best, the runner-up and a count off the list
		// (pe.c:122-160); the same three numbers come out of two scans without the list
		int64_t max = -1;
		uint64_t top1 = 0, top2 = 0;
		int max_i[2] = {-1, -1}, n_sc = 0;
		ps.each([&](int i, int j, int64_t score) {
			if (score > max) max = score, max_i[ps.seg(j)] = j, max_i[ps.seg(i)] = i;
			const uint64_t v = (uint64_t)score;
			if (n_sc == 0) top1 = v;
			else if (v >= top1) top2 = top1, top1 = v;
			else if (n_sc == 1 || v > top2) top2 = v;
			++n_sc;
		});
		if (n_sc > 0 && max > 0) {
			int n_sub = 0;
			ps.each([&](int, int, int64_t score) { if (((uint64_t)score >> 32) + (uint64_t)sub_diff >= (uint64_t)max >> 32) ++n_sub; });
			HitRec *r[2] = {ps.hit(max_i[0]), ps.hit(max_i[1])};
			r[0]->bits |= HB_PROPER, r[1]->bits |= HB_PROPER;
			for (int s = 0; s < 2; ++s) {
				if (r[s]->id != r[s]->parent) { // the paired hit becomes the primary of its family
					HitRec *p = &regs[s][r[s]->parent];
					const int pid = p->id;
					for (int i = 0; i < n_regs[s]; ++i) if (regs[s][i].parent == pid) regs[s][i].parent = r[s]->id;
					p->bits &= ~0xffu;
				}
				if (!(r[s]->bits & HB_SAM_PRI)) {
					for (int i = 0; i < n_regs[s]; ++i) regs[s][i].bits &= ~HB_SAM_PRI;
					r[s]->bits |= HB_SAM_PRI;
				}
			}
			int q0 = (int)HB_MAPQ(r[0]->bits), q1 = (int)HB_MAPQ(r[1]->bits);
			int mapq_pe = q0 > q1 ? q0 : q1;
			if (n_sc > 1) {
				if (n_sub >= lt.n) ok = false;
				else {
					const int alt = (int)(6.02f * (uint64_t)((max >> 32) - (int64_t)(top2 >> 32)) / match_sc - 4.343f * lt.li[n_sub]);
					mapq_pe = mapq_pe < alt ? mapq_pe : alt;
				}
			}
			if (q0 < mapq_pe) q0 = (int)(.2f * q0 + .8f * mapq_pe + .499f);
			if (q1 < mapq_pe) q1 = (int)(.2f * q1 + .8f * mapq_pe + .499f);
			if (n_sc == 1) { if (q0 < 2) q0 = 2; if (q1 < 2) q1 = 2; }
			else if ((uint64_t)max >> 32 > top2 >> 32) { if (q0 < 1) q0 = 1; if (q1 < 1) q1 = 1; }
			r[0]->bits = (r[0]->bits & ~0xffu) | ((uint32_t)q0 & 0xffu);
			r[1]->bits = (r[1]->bits & ~0xffu) | ((uint32_t)q1 & 0xffu);
		}
	}
	{ // mm_set_pe_thru (pe.c:45-63): both mates cover one short fragment end to end
		int n_pri[2] = {0, 0}, pri[2] = {-1, -1};
		for (int s = 0; s < 2; ++s)
			for (int i = 0; i < n_regs[s]; ++i)
				if (regs[s][i].id == regs[s][i].parent) ++n_pri[s], pri[s] = i;
		if (n_pri[0] == 1 && n_pri[1] == 1) {
			HitRec *p = &regs[0][pri[0]], *q = &regs[1][pri[1]];
			const int d1 = p->rs - q->rs, d2 = p->re - q->re;
			if (p->rid == q->rid && ((p->bits ^ q->bits) & HB_REV) == 0 && (d1 < 0 ? -d1 : d1) < 3 && (d2 < 0 ? -d2 : d2) < 3
				&& ((p->qs == 0 && qlens[1] - q->qe == 0) || (q->qs == 0 && qlens[0] - p->qe == 0)))
				p->bits |= HB_PE_THRU, q->bits |= HB_PE_THRU;
		}
	}
	return ok;
}

// ---- warp-cooperative mm_set_parent for fragments with many chains -------------------------------------------------------
// A read pair drawn from a high-copy repeat arrives with thousands of chains.  The hits are still judged one after the other
// (each against the primaries before it), but the two scans over the primaries that the judgement needs become O(1):
//   * the part of hit i no primary covers (hit.c:128-141) is a population count: the query positions covered by ANY primary
//     are kept as a bitmap, 32 positions per lane (a primary that does not overlap hit i covers none of its positions);
//   * the first primary that masks hit i (hit.c:143-161) is found by one ballot over 32 primaries at a time.
// The code is written against a tiny warp interface so that tests/emu/ can run it on the CPU: ballot(f) / sum(f) evaluate f(lane)
// on every lane and combine; each(f) runs f(lane) on every lane; everything else is computed redundantly by all lanes.
#define HIT_COVER_BITS 2048  // query positions the coverage bitmap holds (two words per lane)

MMG_HD uint32_t hit_range_word(int s, int e, int word)
{ // bits of [s, e) that fall into positions [32 word, 32 word + 32)
	const int lo = s - 32 * word, hi = e - 32 * word;
	if (hi <= 0 || lo >= 32) return 0;
	const uint32_t a = lo <= 0 ? 0xffffffffu : 0xffffffffu << lo, b = hi >= 32 ? 0xffffffffu : ~(0xffffffffu << hi);
	return a & b;
}

#ifndef HIT_PRIM_CACHE
#define HIT_PRIM_CACHE 96   // primaries whose query interval and counters are kept in fast memory (shared memory on the device)
#endif
#define HIT_FAST_WORDS (5 * HIT_PRIM_CACHE + 64)   // ints of fast scratch a warp needs: 5 per cached primary + the coverage bitmap

// n hits with 0 <= qs < qe <= HIT_COVER_BITS, no alignment records yet.  w[]: n ints (the primaries, as in hit_set_parent);
// fast[]: HIT_FAST_WORDS ints of fast scratch.  Result == hit_set_parent's.
// Hits are judged 32 at a time, one per lane, against the primaries known so far: what a masked hit does to its primary (a
// maximum and a count) does not depend on the order, so a group of hits that all find a masking primary is committed at once.
// The first hit of a group that finds none becomes a primary; the hits after it are judged again with it in place.
template <class W>
MMG_HDN inline void hit_set_parent_warp(const W &wp, float mask_level, int n, HitRec *r, bool hard_mask_level, int32_t *w, int32_t *fast)
{
	if (n <= 0) return;
	int32_t *pc = fast;                                              // qs, qe, subsc, n_sub, cnt of the first primaries
	uint32_t *cbm = reinterpret_cast<uint32_t*>(fast + 5 * HIT_PRIM_CACHE); // query positions covered by any primary
	wp.each([&](int l) { for (int i = l; i < n; i += 32) r[i].id = i; cbm[l] = hit_range_word(r[0].qs, r[0].qe, l), cbm[l + 32] = hit_range_word(r[0].qs, r[0].qe, l + 32); });
	wp.one([&]() { w[0] = 0, r[0].parent = 0; pc[0] = r[0].qs, pc[1] = r[0].qe, pc[2] = r[0].subsc, pc[3] = r[0].n_sub, pc[4] = r[0].cnt; });
	int k = 1;
	typename W::template Var<int32_t> verdict;
	for (int base = 1; base < n;) {
		wp.each([&](int l) {
			const int i = base + l;
			if (i >= n) { verdict(l) = -2; return; }
			const int si = r[i].qs, ei = r[i].qe;
			int uncov_len = 0, j = -1;
			bool judge = true;
			if (!hard_mask_level) {
				int cov = 0;
				for (int wd = si >> 5; wd <= (ei - 1) >> 5; ++wd) cov += mmg_popc(cbm[wd] & hit_range_word(si, ei, wd));
				if (cov == 0) judge = false;
				else uncov_len = (ei - si) - cov;
			}
			if (judge)
				for (int t = 0; t < k; ++t) {
					int sj, ej;
					if (t < HIT_PRIM_CACHE) sj = pc[5 * t], ej = pc[5 * t + 1];
					else sj = r[w[t]].qs, ej = r[w[t]].qe;
					if (ej <= si || sj >= ei) continue;
					const int mn = ej - sj < ei - si ? ej - sj : ei - si, mx = ej - sj > ei - si ? ej - sj : ei - si;
					const int ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj) : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
					if ((float)ol / mn - (float)uncov_len / mx > mask_level) { j = t; break; }
				}
			verdict(l) = j;
		});
		const unsigned m_new = wp.ballot([&](int l) { return verdict(l) == -1; });
		const int first_new = m_new ? mmg_ffs(m_new) - 1 : 32;
		wp.each([&](int l) { // the hits before the first new primary are masked: commit them
			const int i = base + l, j = verdict(l);
			if (l >= first_new || j < 0) return;
			HitRec *ri = &r[i];
			ri->parent = w[j]; // a primary is its own parent
			if (j < HIT_PRIM_CACHE) { W::amax(&pc[5 * j + 2], ri->score); if (ri->cnt >= pc[5 * j + 4]) W::aadd(&pc[5 * j + 3], 1); }
			else { HitRec *rp = &r[w[j]]; W::amax(&rp->subsc, ri->score); if (ri->cnt >= rp->cnt) W::aadd(&rp->n_sub, 1); }
		});
		if (first_new < 32) {
			const int i = base + first_new, si = r[i].qs, ei = r[i].qe;
			wp.one([&]() {
				w[k] = i, r[i].parent = i, r[i].n_sub = 0;
				if (k < HIT_PRIM_CACHE) pc[5 * k] = si, pc[5 * k + 1] = ei, pc[5 * k + 2] = r[i].subsc, pc[5 * k + 3] = 0, pc[5 * k + 4] = r[i].cnt;
			});
			wp.each([&](int l) { cbm[l] |= hit_range_word(si, ei, l), cbm[l + 32] |= hit_range_word(si, ei, l + 32); });
			++k;
			base = i + 1;
		} else base += 32;
	}
	wp.each([&](int l) { for (int t = l; t < k && t < HIT_PRIM_CACHE; t += 32) r[w[t]].subsc = pc[5 * t + 2], r[w[t]].n_sub = pc[5 * t + 3]; });
}

// pe.c:6-43 for a long hit list, by a warp.  The reference compacts the list in place while it reads the parents' records from
// it, so a hit's verdict can depend on what was kept before it; but almost every hit of a repeat-family fragment is dropped,
// and a group of 32 hits in which nothing is kept changes nothing.  Such groups are decided in one step; any other group runs
// the reference's loop on one lane.  map[]: n ints.
template <class W>
MMG_HDN inline int hit_select_sub_multi_warp(const W &wp, float pri_ratio, float pri1, float pri2, int max_gap_ref, int min_diff, int best_n, int n_segs, const int32_t *qlens,
                                             int n, HitRec *r, int32_t *map, int32_t *state /* 2 ints of fast scratch */)
{
	if (!(pri_ratio > 0.0f && n > 0)) return n;
	const int max_dist = n_segs == 2 ? qlens[0] + qlens[1] + max_gap_ref : 0;
	auto verdict = [&](int i) -> int { // 0: dropped, 1: primary, 2: secondary that passes the score rules (pe.c:17-33)
		if (r[i].parent == i) return 1;
		if (r[i].score + min_diff >= r[r[i].parent].score) return 2;
		const HitRec *p = &r[r[i].parent], *q = &r[i];
		const bool prev = p->bits & HB_REV, qrev = q->bits & HB_REV;
		if (prev == qrev && p->rid == q->rid && q->re - p->rs < max_dist && p->re - q->rs < max_dist) return q->score >= p->score * pri1 ? 2 : 0;
		const int par_both = (n_segs == 2 && p->qs < qlens[0] && p->qe > qlens[0]);
		const int chi_both = (n_segs == 2 && q->qs < qlens[0] && q->qe > qlens[0]);
		if (chi_both || chi_both == par_both) return q->score >= p->score * pri_ratio ? 2 : 0;
		return q->score >= p->score * pri2 ? 2 : 0;
	};
	wp.one([&]() { state[0] = 0, state[1] = 0; }); // k, n_2nd
	for (int base = 0; base < n; base += 32) {
		const int k = state[0], n_2nd = state[1];
		const unsigned m_pri = wp.ballot([&](int l) { return base + l < n && verdict(base + l) == 1; });
		const unsigned m_2nd = wp.ballot([&](int l) { return base + l < n && verdict(base + l) == 2; });
		if (m_pri == 0 && (m_2nd == 0 || n_2nd >= best_n)) { // nothing of this group survives: no record moves
			wp.one([&]() { state[1] = n_2nd + mmg_popc(m_2nd); });
			continue;
		}
		wp.one([&]() {
			int kk = k, n2 = n_2nd;
			const int end = base + 32 < n ? base + 32 : n;
			for (int i = base; i < end; ++i) {
				const int v = verdict(i);
				bool keep = v != 0;
				if (keep && r[i].parent != i && n2++ >= best_n) keep = false;
				if (keep) r[kk++] = r[i];
			}
			state[0] = kk, state[1] = n2;
		});
	}
	const int k = state[0];
	if (k != n) { // mm_sync_regs (hit.c:214-236)
		int max_id = -1;
		for (int i = 0; i < k; ++i) max_id = max_id > r[i].id ? max_id : r[i].id;
		wp.each([&](int l) { for (int i = l; i <= max_id; i += 32) map[i] = -1; });
		wp.one([&]() {
			for (int i = 0; i < k; ++i) if (r[i].id >= 0) map[r[i].id] = i;
			for (int i = 0; i < k; ++i) {
				HitRec *h = &r[i];
				h->id = i;
				if (h->parent == HIT_PARENT_TMP_PRI) h->parent = i;
				else if (h->parent >= 0 && map[h->parent] >= 0) h->parent = map[h->parent];
				else h->parent = HIT_PARENT_UNSET;
			}
			hit_set_sam_pri(k, r);
		});
	}
	return k;
}
