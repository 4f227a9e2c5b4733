// mmg_aln.h -- base-level alignment of a hit for the short-read presets, cut into the two halves a batched DP needs:
//
//   aln_plan_sr()   reads the hit's anchors and the packed sequences, chooses the anchor stretch and the extension windows,
//                   scores the stretch, and emits the (at most three) DP jobs of the region           -> RegionPlan
//   aln_build_sr()  takes the DP results, stitches the CIGAR in place in the shard's word arena, derives the final
//                   coordinates, cleans the CIGAR up and re-scores it; may cut the hit at a z-drop     -> HitRec (+ tail hit)
//
// Nothing here allocates: sequences are read as 4-bit codes straight from the resident read batch and the index, the CIGAR is
// written where a prefix sum reserved room for it.  __host__ __device__ (tests/emu/ runs it against mm_map_frag on the CPU).
//
// Reference behaviour (src/minimap2-master_remapping/): mm_align1 align.c:565-788 (sr branches), mm_max_stretch :495-521,
// mm_test_zdrop :47-89, mm_append_cigar :288-311, mm_fix_cigar :91-167, mm_update_extra :240-286, mm_split_reg hit.c:90-107.
#pragma once
#include "mmg_hits.h"

#define ALN_EZ_RIGHT      0x02
#define ALN_EZ_EXTZ_ONLY  0x40
#define ALN_EZ_REV_CIGAR  0x80

struct SeqView {            // one read of the resident batch + the index sequences
	const uint32_t *Q;      // packed reads (mates already in mapping orientation)
	uint64_t q_off;         // first base of this read in Q
	int32_t qlen;
	const uint32_t *S;      // packed reference
	const uint64_t *ref_off;
	const uint32_t *ref_len;
};

// qseq0[rev][i] of mm_align_skeleton (align.c:865-870)
MMG_HD int aln_q(const SeqView &v, int rev, int i)
{
	if (!rev) return mmg_seq4_get(v.Q, v.q_off + (uint64_t)i);
	const int c = mmg_seq4_get(v.Q, v.q_off + (uint64_t)(v.qlen - 1 - i));
	return c < 4 ? 3 - c : 4;
}
MMG_HD int aln_t(const SeqView &v, int rid, int pos) { return mmg_seq4_get(v.S, v.ref_off[rid] + (uint64_t)pos); }

// ksw_gen_simple_mat (align.c:9-22) entry for target code ct, query code cq
MMG_HD int aln_mat(const HitOpt &o, int ct, int cq)
{
	const int amb = o.sc_ambi > 0 ? -o.sc_ambi : o.sc_ambi;
	if (ct > 3 || cq > 3) return amb;
	return ct == cq ? (o.a < 0 ? -o.a : o.a) : (o.b > 0 ? -o.b : o.b);
}

struct DpJob {              // == mmg_ksw_job_t (include/mmg.h)
	int32_t seq_id, q_rev, q_start, q_len, rid, t_start, t_len, reversed, w, zdrop, end_bonus, flag;
};
struct DpRes { KswEz ez; uint64_t cigar_off; };   // == the device result record of mmg_ksw.cu
struct DpView { const DpRes *res; const uint32_t *cig; };

struct RegionPlan {
	int32_t rs, qs, re, qe;          // the anchor stretch the alignment is pinned to, on the hit's strand
	int32_t rs0, re0;                // reach of the two extensions on the reference (query: 0 and qlen)
	int32_t as1, cnt1;               // the stretch's anchors
	int32_t fill_score, zdrop_code;  // ungapped score of the stretch; 1: it z-drops and gets a DP pass of its own
	int32_t job_left, job_fill, job_right;   // job ids; -1: none; -2: refused (matrix larger than max_sw_mat), counts as z-dropped
	int32_t state;                   // 0: to be planned, 1: jobs in flight, 2: aligned
};

// the longest run of anchors on one diagonal, by summed span (align.c:495-521)
MMG_HD void aln_max_stretch(const HitRec &r, const mm128 *a, int32_t *as1, int32_t *cnt1)
{
	*as1 = r.as, *cnt1 = r.cnt;
	if (r.cnt < 2) return;
	int32_t best = -1, best_i = -1, best_len = 0, run = hit_span(a[r.as]), len = 1, i;
	for (i = r.as + 1; i < r.as + r.cnt; ++i) {
		const int32_t sp = hit_span(a[i]), dr = hit_rpos(a[i]) - hit_rpos(a[i - 1]), dq = hit_qpos(a[i]) - hit_qpos(a[i - 1]);
		if (dq == dr) run += dq < sp ? dq : sp, ++len;
		else {
			if (run > best) best = run, best_len = len, best_i = i - len;
			run = sp, len = 1;
		}
	}
	if (run > best) best_len = len, best_i = i - len;
	*as1 = best_i, *cnt1 = best_len;
}

// how far an extension may reach beyond l unaligned query bases (align.c:614-621)
MMG_HD int32_t aln_reach(const HitOpt &o, int32_t l)
{
	return l + (l * o.a + o.end_bonus > o.q ? (l * o.a + o.end_bonus - o.q) / o.e : 0);
}

// job or its refusal (mm_align_pair, align.c:313-327): empty slices give the reset result, oversized ones count as z-dropped
template <class Emit>
MMG_HD int32_t aln_job(const HitOpt &o, Emit &emit, int read_id, int rev, int q0, int ql, int rid, int t0, int tl, int reversed, int bw, int end_bonus, int zdrop, int flag)
{
	if (o.max_sw_mat > 0 && (int64_t)tl * ql > o.max_sw_mat) return -2;
	if (ql <= 0 || tl <= 0) return -1;
	DpJob j;
	j.seq_id = read_id, j.q_rev = rev, j.q_start = q0, j.q_len = ql, j.rid = rid, j.t_start = t0, j.t_len = tl, j.reversed = reversed;
	j.w = bw, j.zdrop = zdrop, j.end_bonus = end_bonus, j.flag = flag;
	return emit(j);
}

template <class Emit>
MMG_HDN inline void aln_plan_sr(const HitOpt &o, const SeqView &v, int read_id, const HitRec &r, const mm128 *a, RegionPlan *pl, Emit &&emit)
{
	pl->job_left = pl->job_fill = pl->job_right = -1;
	pl->fill_score = 0, pl->zdrop_code = 0, pl->state = 1;
	if (r.cnt == 0) { pl->cnt1 = 0; return; }
	const int rev = (r.bits & HB_REV) ? 1 : 0, rid = (int32_t)(a[r.as].x << 1 >> 33), qlen = v.qlen;
	const int32_t ref_len = (int32_t)v.ref_len[rid], bw = (int)(o.bw * 1.5 + 1.);
	int32_t as1, cnt1;
	aln_max_stretch(r, a, &as1, &cnt1);
	const mm128 fa = a[as1], la = a[as1 + cnt1 - 1];
	const int32_t rs = hit_rpos(fa) + 1 - hit_span(fa), qs = hit_qpos(fa) + 1 - hit_span(fa), re = hit_rpos(la) + 1, qe = hit_qpos(la) + 1;
	int32_t rs0 = rs - aln_reach(o, qs), re0 = re + aln_reach(o, qlen - qe);
	if (rs0 < 0) rs0 = 0;
	if (re0 > ref_len) re0 = ref_len;
	pl->rs = rs, pl->qs = qs, pl->re = re, pl->qe = qe, pl->rs0 = rs0, pl->re0 = re0, pl->as1 = as1, pl->cnt1 = cnt1;
	// the stretch itself is taken as an ungapped match run; an N on either side scores +e2 (align.c:724-731), and the
	// z-drop test walks the same run with the substitution matrix (align.c:47-76; the inversion part is off for sr)
	{
		int32_t score = 0, zs = 0, zmax = INT32_MIN, max_zdrop = 0;
		for (int32_t j = 0; j < qe - qs; ++j) {
			const int cq = aln_q(v, rev, qs + j), ct = aln_t(v, rid, rs + j);
			score += (cq >= 4 || ct >= 4) ? o.e2 : cq == ct ? o.a : -o.b;
			zs += aln_mat(o, ct, cq);
			if (zs < zmax) { if (zmax - zs > max_zdrop) max_zdrop = zmax - zs; }
			else zmax = zs;
		}
		pl->fill_score = score, pl->zdrop_code = max_zdrop > o.zdrop ? 1 : 0;
	}
	if (qs > 0 && rs > 0) // left extension: both slices reversed, gaps right-aligned, CIGAR reversed (align.c:690-705)
		pl->job_left = aln_job(o, emit, read_id, rev, 0, qs, rid, rs0, rs - rs0, 1, bw, o.end_bonus, (r.bits & HB_SPLIT_INV) ? o.zdrop_inv : o.zdrop,
		                       ALN_EZ_EXTZ_ONLY | ALN_EZ_RIGHT | ALN_EZ_REV_CIGAR);
	if (pl->zdrop_code) pl->job_fill = aln_job(o, emit, read_id, rev, qs, qe - qs, rid, rs, re - rs, 0, bw, -1, o.zdrop, 0);
	if (qe < qlen && re < re0) // right extension (align.c:760-771); dropped later if the stretch z-drops
		pl->job_right = aln_job(o, emit, read_id, rev, qe, qlen - qe, rid, re, re0 - re, 0, bw, o.end_bonus, o.zdrop, ALN_EZ_EXTZ_ONLY);
}

struct AlnEz { int32_t max, zdropped, max_q, max_t, mqe_t, score, n_cigar, reach_end; const uint32_t *cigar; };
MMG_HD AlnEz aln_result(const DpView &dp, int32_t job)
{
	AlnEz z;
	z.max = 0, z.zdropped = job == -2 ? 1 : 0, z.max_q = z.max_t = z.mqe_t = -1, z.score = MMG_KSW_NEG_INF, z.n_cigar = 0, z.reach_end = 0, z.cigar = nullptr;
	if (job >= 0) {
		const DpRes &d = dp.res[job];
		z.max = (int32_t)d.ez.max, z.zdropped = d.ez.zdropped, z.max_q = d.ez.max_q, z.max_t = d.ez.max_t, z.mqe_t = d.ez.mqe_t, z.score = d.ez.score;
		z.n_cigar = d.ez.n_cigar, z.reach_end = d.ez.reach_end, z.cigar = dp.cig + d.cigar_off;
	}
	return z;
}

// words a region's alignment record needs: header + every CIGAR operation its DP jobs returned + the stretch
MMG_HD uint32_t aln_record_words(const RegionPlan &pl, const DpView &dp)
{
	if (pl.cnt1 == 0) return 0;
	uint32_t n = 1;
	if (pl.job_left >= 0) n += (uint32_t)dp.res[pl.job_left].ez.n_cigar;
	if (pl.job_fill >= 0) n += (uint32_t)dp.res[pl.job_fill].ez.n_cigar;
	if (pl.job_right >= 0) n += (uint32_t)dp.res[pl.job_right].ez.n_cigar;
	return n + (uint32_t)(sizeof(HitExtra) / 4);
}

MMG_HD uint32_t aln_roundup32(uint32_t x) { --x; x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16; return ++x; }

// mm_append_cigar (align.c:288-311) on a record whose room is already reserved; `capacity` follows the reference's growth rule
MMG_HD void aln_append(HitExtra *p, uint32_t *cigar, bool *fresh, uint32_t n, const uint32_t *src)
{
	if (n == 0) return;
	if (*fresh) p->capacity = aln_roundup32(n + (uint32_t)(sizeof(HitExtra) / 4)), *fresh = false;
	else if (p->n_cigar + n + sizeof(HitExtra) / 4 > p->capacity) p->capacity = aln_roundup32(p->n_cigar + n + (uint32_t)(sizeof(HitExtra) / 4));
	uint32_t k = 0;
	if (p->n_cigar > 0 && (cigar[p->n_cigar - 1] & 0xf) == (src[0] & 0xf)) cigar[p->n_cigar - 1] += src[0] >> 4 << 4, k = 1;
	for (; k < n; ++k) cigar[p->n_cigar++] = src[k];
}

// mm_fix_cigar (align.c:91-167): indels move as far left as the sequences allow, runs of I and D collapse to one of each,
// emptied operations go, and a leading I or D is cut off the alignment.  q(i) / t(i): bases from the alignment's start.
template <class QF, class TF>
MMG_HDN inline void aln_fix_cigar(HitRec *r, HitExtra *p, uint32_t *cg, QF &&q, TF &&t, int32_t *qshift, int32_t *tshift)
{
	*qshift = *tshift = 0;
	if (p->n_cigar <= 1) return;
	int32_t toff = 0, qoff = 0;
	bool shrink = false;
	for (uint32_t k = 0; k < p->n_cigar; ++k) {
		const uint32_t op = cg[k] & 0xf, len = cg[k] >> 4;
		if (len == 0) shrink = true;
		if (op == 0) toff += len, qoff += len;
		else if (op == 1 || op == 2) {
			if (k > 0 && k < p->n_cigar - 1 && (cg[k - 1] & 0xf) == 0 && (cg[k + 1] & 0xf) == 0) {
				const int32_t prev_len = (int32_t)(cg[k - 1] >> 4), o = op == 1 ? qoff : toff;
				int32_t l = 0;
				if (op == 1) { for (; l < prev_len; ++l) if (q(o - 1 - l) != q(o + (int32_t)len - 1 - l)) break; }
				else { for (; l < prev_len; ++l) if (t(o - 1 - l) != t(o + (int32_t)len - 1 - l)) break; }
				if (l > 0) cg[k - 1] -= (uint32_t)l << 4, cg[k + 1] += (uint32_t)l << 4, qoff -= l, toff -= l;
				if (l == prev_len) shrink = true;
			}
			if (op == 1) qoff += len; else toff += len;
		} else if (op == 3) toff += len;
	}
	for (uint32_t k = 0; k + 2 < p->n_cigar; ++k)
		if ((cg[k] & 0xf) > 0 && (cg[k] & 0xf) + (cg[k + 1] & 0xf) == 3) {
			uint32_t l, s[3] = {0, 0, 0};
			for (l = k; l < p->n_cigar; ++l) {
				const uint32_t op = cg[l] & 0xf;
				if (op == 1 || op == 2 || cg[l] >> 4 == 0) s[op] += cg[l] >> 4;
				else break;
			}
			if (s[1] > 0 && s[2] > 0 && l - k > 2) {
				cg[k] = s[1] << 4 | 1, cg[k + 1] = s[2] << 4 | 2;
				for (k += 2; k < l; ++k) cg[k] &= 0xf;
				shrink = true;
			}
			k = l;
		}
	if (shrink) {
		uint32_t l = 0;
		for (uint32_t k = 0; k < p->n_cigar; ++k) if (cg[k] >> 4 != 0) cg[l++] = cg[k];
		p->n_cigar = l;
		l = 0;
		for (uint32_t k = 0; k < p->n_cigar; ++k)
			if (k == p->n_cigar - 1 || (cg[k] & 0xf) != (cg[k + 1] & 0xf)) cg[l++] = cg[k];
			else cg[k + 1] += cg[k] >> 4 << 4;
		p->n_cigar = l;
	}
	if ((cg[0] & 0xf) == 1 || (cg[0] & 0xf) == 2) {
		const int32_t l = (int32_t)(cg[0] >> 4);
		if ((cg[0] & 0xf) == 1) {
			if (r->bits & HB_REV) r->qe -= l; else r->qs += l;
			*qshift = l;
		} else r->rs += l, *tshift = l;
		--p->n_cigar;
		for (uint32_t k = 0; k < p->n_cigar; ++k) cg[k] = cg[k + 1];
	}
}

// mm_update_extra (align.c:240-286) without the =/X rewrite (--eqx keeps the host path)
template <class QF, class TF>
MMG_HDN inline void aln_update_extra(const HitOpt &o, HitRec *r, HitExtra *p, uint32_t *cg, QF &&q0, TF &&t0)
{
	int32_t qshift, tshift;
	aln_fix_cigar(r, p, cg, q0, t0, &qshift, &tshift);
	int32_t s = 0, max = 0, toff = tshift, qoff = qshift, blen = 0, mlen = 0;
	uint32_t n_ambi_tot = p->n_ambi_strand & 0x3fffffffu;
	for (uint32_t k = 0; k < p->n_cigar; ++k) {
		const uint32_t op = cg[k] & 0xf, len = cg[k] >> 4;
		if (op == 0) {
			int n_ambi = 0, n_diff = 0;
			for (uint32_t l = 0; l < len; ++l) {
				const int cq = q0(qoff + (int32_t)l), ct = t0(toff + (int32_t)l);
				if (ct > 3 || cq > 3) ++n_ambi;
				else if (ct != cq) ++n_diff;
				s += aln_mat(o, ct, cq);
				if (s < 0) s = 0;
				else max = max > s ? max : s;
			}
			blen += (int32_t)len - n_ambi, mlen += (int32_t)len - (n_ambi + n_diff), n_ambi_tot += (uint32_t)n_ambi;
			toff += len, qoff += len;
		} else if (op == 1 || op == 2) {
			int n_ambi = 0;
			for (uint32_t l = 0; l < len; ++l) if ((op == 1 ? q0(qoff + (int32_t)l) : t0(toff + (int32_t)l)) > 3) ++n_ambi;
			blen += (int32_t)len - n_ambi, n_ambi_tot += (uint32_t)n_ambi;
			s -= o.q + o.e * (int32_t)len;
			if (s < 0) s = 0;
			if (op == 1) qoff += len; else toff += len;
		} else if (op == 3) toff += len;
	}
	r->blen = blen, r->mlen = mlen;
	p->n_ambi_strand = (p->n_ambi_strand & 0xc0000000u) | (n_ambi_tot & 0x3fffffffu);
	p->dp_max = max;
}

// mm_split_reg (hit.c:90-107): the hit keeps its first n anchors, the rest becomes a hit of its own
MMG_HDN inline void aln_split(HitRec *r, HitRec *r2, int n, int qlen, const mm128 *a)
{
	if (n <= 0 || n >= r->cnt) return;
	*r2 = *r;
	r2->id = -1, r2->p = 0;
	r2->bits &= ~(HB_SAM_PRI | HB_SPLIT_INV);
	r2->cnt = r->cnt - n;
	r2->score = (int32_t)(r->score * ((float)r2->cnt / r->cnt) + .499);
	r2->as = r->as + n;
	if (r->parent == r->id) r2->parent = HIT_PARENT_TMP_PRI;
	hit_set_coor(r2, qlen, a);
	r->cnt -= r2->cnt, r->score -= r2->score;
	hit_set_coor(r, qlen, a);
	r->bits |= 1u << 8, r2->bits |= 2u << 8;
}

// The second half of mm_align1 for a planned region.  xw + xoff: the room reserved for this hit's alignment record
// (aln_record_words() words).  Returns 1 when a z-drop cut produced a tail hit in *r2.
MMG_HDN inline int aln_build_sr(const HitOpt &o, const SeqView &v, HitRec *r, HitRec *r2, const mm128 *a, const RegionPlan &pl, const DpView &dp, uint32_t *xw, uint64_t xoff)
{
	r2->cnt = 0;
	if (r->cnt == 0 || pl.cnt1 == 0) return 0;
	const int rev = (r->bits & HB_REV) ? 1 : 0, rid = (int32_t)(a[r->as].x << 1 >> 33), qlen = v.qlen;
	HitExtra *p = reinterpret_cast<HitExtra*>(xw + xoff);
	uint32_t *cg = xw + xoff + sizeof(HitExtra) / 4;
	bool fresh = true, dropped = false;
	p->capacity = 0, p->dp_score = 0, p->dp_max = 0, p->dp_max2 = 0, p->n_ambi_strand = 0, p->n_cigar = 0;
	int32_t rs1 = pl.rs, qs1 = pl.qs, re1, qe1;
	if (pl.qs > 0 && pl.rs > 0) {
		const AlnEz ez = aln_result(dp, pl.job_left);
		if (ez.n_cigar > 0) { aln_append(p, cg, &fresh, (uint32_t)ez.n_cigar, ez.cigar); p->dp_score += ez.max; }
		rs1 = pl.rs - (ez.reach_end ? ez.mqe_t + 1 : ez.max_t + 1);
		qs1 = pl.qs - (ez.reach_end ? pl.qs : ez.max_q + 1);
	}
	re1 = pl.re, qe1 = pl.qe;
	{ // the stretch: an ungapped run, or the DP pass a z-dropping run gets (align.c:709-758 with i = cnt1 - 1)
		AlnEz ez;
		const uint32_t one = (uint32_t)(pl.qe - pl.qs) << 4;
		if (pl.zdrop_code) ez = aln_result(dp, pl.job_fill);
		else { ez = aln_result(dp, -1); ez.score = pl.fill_score, ez.n_cigar = 1, ez.cigar = &one; }
		if (ez.n_cigar > 0) aln_append(p, cg, &fresh, (uint32_t)ez.n_cigar, ez.cigar);
		if (ez.zdropped) { // cut the hit behind the last anchor the alignment still reaches (align.c:741-754)
			int j;
			for (j = pl.cnt1 - 2; j >= 0; --j) if (hit_rpos(a[pl.as1 + j]) <= pl.rs + ez.max_t) break;
			dropped = true;
			if (j < 0) j = 0;
			p->dp_score += ez.max;
			re1 = pl.rs + (ez.max_t + 1), qe1 = pl.qs + (ez.max_q + 1);
			if (pl.cnt1 - (j + 1) >= o.min_cnt) aln_split(r, r2, pl.as1 + j + 1 - r->as, qlen, a);
		} else p->dp_score += ez.score;
	}
	if (!dropped && pl.qe < qlen && pl.re < pl.re0) {
		const AlnEz ez = aln_result(dp, pl.job_right);
		if (ez.n_cigar > 0) { aln_append(p, cg, &fresh, (uint32_t)ez.n_cigar, ez.cigar); p->dp_score += ez.max; }
		re1 = pl.re + (ez.reach_end ? ez.mqe_t + 1 : ez.max_t + 1);
		qe1 = pl.qe + (ez.reach_end ? qlen - pl.qe : ez.max_q + 1);
	}
	r->rs = rs1, r->re = re1;
	if (rev) r->qs = qlen - qe1, r->qe = qlen - qs1;
	else r->qs = qs1, r->qe = qe1;
	if (!fresh) {
		r->p = xoff + 1;
		aln_update_extra(o, r, p, cg, [&](int32_t i) { return aln_q(v, rev, qs1 + i); }, [&](int32_t i) { return aln_t(v, rid, rs1 + i); });
	}
	return r2->cnt > 0 ? 1 : 0;
}
