// mmg_kswfast2.h -- K4, fast form (band never clips), one thread per job, TWO cells per instruction.
//
// mmg_ksw_fast_run (mmg_core.h) spends ~105 scalar instructions on a cell: six sign-extending unpacks and six packs around ~25
// instructions of arithmetic.  Here the thread walks an anti-diagonal in PAIRS of cells (t even, t + 1) with the pair function of
// the literal 16x2 kernel (mmg_kswdpx_pair: VIADD.16x2 / VIMNMX3.S16x2 / VIADDMNMX.S16x2, direction from one max over
// value * 8 + preference), without its int8 folds: in this form no stale lane is ever read, and a valid cell never leaves the
// int8 range (checked by the emulation tests through MMG_KSW_RANGE on the scalar form).  The state of a pair is one 16-byte slot
// {u0 u1 v0 v1 | x0 x1 y0 y1 | x2.. y2.. | target bases} in a circular window of (min(qlen, tlen) >> 1) + 3 slots -- one LDS.128 and
// one STS.128 per two cells.
//
// Pairs are aligned on the target coordinate, anti-diagonals are not: when st0 is odd the low half of the first pair is cell
// st0 - 1, when en0 is even the high half of the last pair is cell en0 + 1.  Both are outside the matrix on this anti-diagonal; what
// the pair function computes for them is garbage that no valid cell ever reads: cell st0 - 1 never becomes valid again (st0 only
// grows), and cell en0 + 1 is either beyond the target or the cell that ENTERS on the next anti-diagonal, whose boundary state is
// written then (ksw2_extd2_sse.c:152-156).  The exact-max scan, the traceback walk and the H bookkeeping touch valid cells only.
#ifndef MMG_KSWFAST2_H
#define MMG_KSWFAST2_H
#include "mmg_kswdpx.h"

MMG_HD int mmg_ksw_fast2_wp(int qlen, int tlen) { return ((qlen < tlen ? qlen : tlen) >> 1) + 3; }      // pair slots of the window
MMG_HD int mmg_ksw_fast2_pw(int qlen, int tlen) { return ((qlen < tlen ? qlen : tlen) + 2 + 3) & ~3; }  // traceback row: cells st0 & ~1 ..
MMG_HD size_t mmg_ksw_fast2_p_bytes(int qlen, int tlen) { return ((size_t)(qlen + tlen - 1) * mmg_ksw_fast2_pw(qlen, tlen) + 15) & ~(size_t)15; }

struct alignas(16) KswSlot { uint32_t w0, w1, w2, w3; };

MMG_HD KswSlot mmg_ksw_slot_ld(const KswSlot *p)
{
#ifdef __CUDA_ARCH__
	const uint4 v = *reinterpret_cast<const uint4*>(p);
	KswSlot s; s.w0 = v.x, s.w1 = v.y, s.w2 = v.z, s.w3 = v.w;
	return s;
#else
	return *p;
#endif
}
MMG_HD void mmg_ksw_slot_st(KswSlot *p, const KswSlot &s)
{
#ifdef __CUDA_ARCH__
	*reinterpret_cast<uint4*>(p) = make_uint4(s.w0, s.w1, s.w2, s.w3);
#else
	*p = s;
#endif
}

// preference of cell k = t - st0 among equal scores of an anti-diagonal (mmg_ksw_max_rank without the en0 cell, which the caller
// handles): the four strided lanes of the reference's scan, each keeping its first hit, merged lane 0..3, then the tail
MMG_HD int32_t mmg_ksw_fast2_pref(int k, int en1k)
{
	const int32_t rk = k < en1k ? (((k >> 2) | (k << 8)) & 0x3ff) : (1024 + (k - en1k));
	return 2046 - rk;
}

// S: window of mmg_ksw_fast2_wp() slots, slot i of this job at S[i * stride]; tb(t) = target base t (asked once per cell, when
// it enters: the kernel reads it from the packed reference instead of keeping a copy in shared memory); qb[(j + 1) * stride]
// query bases with one padding element before and after; p: rows of mmg_ksw_fast2_pw() bytes
template <int kMode, class TB>
MMG_HDN inline void mmg_ksw_fast2_run(const KswGeom &g, int flag, int zdrop, KswSlot *S, TB &&tb, const uint8_t *qb, int stride,
                                      uint32_t *p, KswEz &ez)
{
	const int qlen = g.qlen, tlen = g.tlen;
	const int Wp = mmg_ksw_fast2_wp(qlen, tlen), PWw = mmg_ksw_fast2_pw(qlen, tlen) >> 2;
	const bool approx = (flag & MMG_EZ_APPROX_MAX) != 0;
	const int32_t n1 = -g.q - g.e, n2 = -g.q2 - g.e2;
	const KswDpxConst cst = mmg_kswdpx_const<kMode>(g);
	const uint8_t *qbp = qb + stride; // qbp[j * stride], j = -1 .. qlen
	int P0 = 0, s0 = 0;               // first pair of the anti-diagonal and its window slot
	int32_t Hst = 0, H0 = 0, last_H0_t = 0;
	uint32_t tb_next = tb(0); // base of the cell that enters next: asked for one anti-diagonal ahead, so that a load from the packed reference is never waited for
#define KF2_SLOT(P_) (S + (size_t)((s0 + ((P_) - P0)) >= Wp ? (s0 + ((P_) - P0)) - Wp : (s0 + ((P_) - P0))) * stride)
#define KF2_U(t_) ((int32_t)(int8_t)(KF2_SLOT((t_) >> 1)->w0 >> (((t_) & 1) << 3)))
#define KF2_V(t_) ((int32_t)(int8_t)(KF2_SLOT((t_) >> 1)->w0 >> (16 + (((t_) & 1) << 3))))
	for (int r = 0; r < qlen + tlen - 1; ++r) {
		const int st0 = r - qlen + 1 > 0 ? r - qlen + 1 : 0, en0 = r < tlen - 1 ? r : tlen - 1, n = en0 - st0 + 1;
		const int32_t fc = mmg_ksw_first_col(g, r);
		if ((st0 >> 1) != P0) { ++P0; if (++s0 == Wp) s0 = 0; } // st0 grows by at most one per anti-diagonal
		uint32_t prev;
		if (st0 == 0) prev = mmg_kswdpx_carry_of(n1, fc, n2);
		else {
			const KswSlot o = mmg_ksw_slot_ld(S + (size_t)(s0 == 0 ? Wp - 1 : s0 - 1) * stride); // pair P0 - 1 (its high half is cell st0 - 1 when st0 is even)
			const KswPair op = {o.w0, o.w1, o.w2};
			prev = mmg_kswdpx_carry(op);
		}
		if (r < tlen) { // the cell entering on the first row: boundary u / y / y2 (ksw2_extd2_sse.c:152-156) and its target base
			KswSlot *sp = KF2_SLOT(r >> 1);
			KswSlot o = mmg_ksw_slot_ld(sp);
			const uint32_t sh = (r & 1) << 3, m8 = 0xffu << sh, m8h = m8 << 16;
			o.w0 = (o.w0 & ~m8) | ((uint32_t)(fc & 0xff) << sh);                         // u
			o.w1 = (o.w1 & ~m8h) | ((uint32_t)(n1 & 0xff) << (16 + sh));                 // y
			o.w2 = (o.w2 & ~m8h) | ((uint32_t)(n2 & 0xff) << (16 + sh));                 // y2
			o.w3 = (o.w3 & ~m8) | (tb_next << sh);
			mmg_ksw_slot_st(sp, o);
			if (r + 1 < tlen) tb_next = tb(r + 1);
		}
		const int np = (en0 >> 1) - P0 + 1;
		const int en1k = (n - 1) / 4 * 4;
		int32_t E = 0, best = INT32_MIN, v_last = 0;
		uint32_t acc = 0;
		uint32_t *prow = p + (size_t)r * PWw;
		int slot = s0;
		const uint8_t *qp = qbp + (ptrdiff_t)(r - 2 * P0) * stride; // query base of the pair's low cell; the high cell's is one before
		int k = 2 * P0 - st0;                                      // k of the pair's low cell: -1 or 0 for the first pair
		for (int pp = 0; pp < np; ++pp) {
			KswSlot *sp = S + (size_t)slot * stride;
			const KswSlot o = mmg_ksw_slot_ld(sp);
			const uint32_t tb0 = o.w3 & 0xff, tb1 = o.w3 >> 8 & 0xff, q0 = qp[0], q1 = *(qp - stride);
			int32_t sc0 = tb0 == q0 ? g.sc_mch : g.sc_mis, sc1 = tb1 == q1 ? g.sc_mch : g.sc_mis;
			if ((tb0 | q0) & 4) sc0 = g.sc_N; // codes are 0..4: either one is N
			if ((tb1 | q1) & 4) sc1 = g.sc_N;
			const KswPair op = {o.w0, o.w1, o.w2};
			uint32_t d2 = 0;
			const KswPair nw = mmg_kswdpx_pair<kMode, false>(cst, op, (uint32_t)(sc0 & 0xff) | (uint32_t)(sc1 & 0xff) << 8, prev, &d2);
			prev = mmg_kswdpx_carry(op);
			KswSlot ns; ns.w0 = nw.w0, ns.w1 = nw.w1, ns.w2 = nw.w2, ns.w3 = o.w3;
			mmg_ksw_slot_st(sp, ns);
			if (kMode) {
				acc = acc >> 16 | d2 << 16;
				if (pp & 1) prow[pp >> 1] = acc;
			}
			if (!approx) { // exact max: E_k = H(r, st0 + k) - H(r, st0) + u_0 accumulates from u - v of neighbouring cells; key = E << 11 | preference
				const int32_t u0 = (int8_t)nw.w0, u1 = (int8_t)(nw.w0 >> 8), v0 = (int8_t)(nw.w0 >> 16), v1 = (int8_t)(nw.w0 >> 24);
				if (k >= 0) {
					E += u0;
					const int32_t key = E * 2048 + mmg_ksw_fast2_pref(k, en1k);
					best = best > key ? best : key;
					E -= v0; v_last = v0;
				}
				if (k + 1 < n) {
					E += u1;
					const int32_t key = E * 2048 + mmg_ksw_fast2_pref(k + 1, en1k);
					best = best > key ? best : key;
					E -= v1; v_last = v1;
				}
			}
			k += 2;
			qp -= 2 * (ptrdiff_t)stride;
			if (++slot == Wp) slot = 0;
		}
		if (kMode && (np & 1)) prow[np >> 1] = acc >> 16;
		if (!approx) {
			const int32_t u0 = KF2_U(st0), v0 = KF2_V(st0);
			const int32_t H_st0 = r == 0 ? v0 - g.qe_pre : Hst + (st0 > 0 ? u0 : v0);
			const int32_t off = H_st0 - u0, H_en0 = off + E + v_last;
			int32_t bh, bt;
			{ // the cell en0 wins every tie (ksw2_extd2_sse.c:319-349 starts from it)
				const int32_t key = (E + v_last) * 2048 + 2047;
				best = best > key ? best : key;
				const int32_t low = best & 2047;
				bh = ((best - low) >> 11) + off; // exact: best - low is a multiple of 2048
				if (low == 2047) bt = en0;
				else { const int32_t rk = 2046 - low; bt = rk >= 1024 ? st0 + en1k + (rk - 1024) : st0 + ((rk & 255) << 2) + (rk >> 8); }
			}
			Hst = H_st0;
			if (en0 == tlen - 1 && H_en0 > ez.mte) ez.mte = H_en0, ez.mte_q = r - ((en0 + 16) / 16 * 16 - 1); // the widened end (ksw2_extd2_sse.c:352)
			if (r - st0 == qlen - 1 && H_st0 > ez.mqe) ez.mqe = H_st0, ez.mqe_t = st0;
			if (mmg_ksw_zdrop(&ez, bh, r, bt, zdrop, g.e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H_en0;
		} else { // ksw2_extd2_sse.c:359-375, reading the cells just written
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					const int32_t d0 = KF2_V(last_H0_t), d1 = KF2_U(last_H0_t + 1);
					if (d0 > d1) H0 += d0; else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) H0 += KF2_V(last_H0_t);
				else { ++last_H0_t; H0 += KF2_U(last_H0_t); }
			} else H0 = KF2_V(0) - g.qe_pre, last_H0_t = 0;
			if ((flag & MMG_EZ_APPROX_DROP) && mmg_ksw_zdrop(&ez, H0, r, last_H0_t, zdrop, g.e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H0;
		}
	}
#undef KF2_SLOT
#undef KF2_U
#undef KF2_V
}

MMG_HDN inline int mmg_ksw_backtrack_fast2(const KswGeom &g, int is_rev, const uint8_t *p, int i0, int j0, uint32_t *cigar)
{ // ksw_backtrack (ksw2.h:119-151) over the pair form's rows; the walk cannot leave the band, so no forced states
	int n = 0, i = i0, j = j0, state = 0;
	const int pw = mmg_ksw_fast2_pw(g.qlen, g.tlen);
#define MMG_PUSH(op, len) do { if (n == 0 || (uint32_t)(op) != (cigar[n - 1] & 0xf)) cigar[n++] = (uint32_t)(len) << 4 | (uint32_t)(op); else cigar[n - 1] += (uint32_t)(len) << 4; } while (0)
	while (i >= 0 && j >= 0) {
		const int r = i + j, st0 = r - g.qlen + 1 > 0 ? r - g.qlen + 1 : 0;
		const uint32_t tmp = p[(size_t)r * pw + i - (st0 & ~1)];
		if (state == 0) state = tmp & 7;
		else if (!(tmp >> (state + 2) & 1)) state = 0;
		if (state == 0) state = tmp & 7;
		if (state == 0) { MMG_PUSH(0, 1); --i, --j; }
		else if (state == 1 || state == 3) { MMG_PUSH(2, 1); --i; }
		else { MMG_PUSH(1, 1); --j; }
	}
	if (i >= 0) MMG_PUSH(2, i + 1);
	if (j >= 0) MMG_PUSH(1, j + 1);
#undef MMG_PUSH
	if (!is_rev)
		for (int k = 0; k < n >> 1; ++k) { uint32_t t = cigar[k]; cigar[k] = cigar[n - 1 - k]; cigar[n - 1 - k] = t; }
	return n;
}

template <class TB>
MMG_HDN inline void mmg_ksw_fast2(const KswGeom &g, int flag, int zdrop, int end_bonus, KswSlot *S, TB &&tb, const uint8_t *qb, int stride,
                                  uint32_t *p, KswEz *ez_out, uint32_t *cigar)
{
	KswEz ez;
	mmg_ksw_reset(&ez);
	const bool with_cigar = !(flag & MMG_EZ_SCORE_ONLY);
	if (!with_cigar) mmg_ksw_fast2_run<0>(g, flag, zdrop, S, tb, qb, stride, p, ez);
	else if (!(flag & MMG_EZ_RIGHT)) mmg_ksw_fast2_run<1>(g, flag, zdrop, S, tb, qb, stride, p, ez);
	else mmg_ksw_fast2_run<2>(g, flag, zdrop, S, tb, qb, stride, p, ez);
	int i0, j0;
	if (with_cigar && mmg_ksw_trace_start(g, flag, end_bonus, &ez, &i0, &j0))
		ez.n_cigar = mmg_ksw_backtrack_fast2(g, !!(flag & MMG_EZ_REV_CIGAR), reinterpret_cast<const uint8_t*>(p), i0, j0, cigar);
	*ez_out = ez;
}
#endif
