// mmg_kswdpx.h -- the literal form of ksw_extd2 (ksw2_extd2_sse.c:184-313) on Blackwell's 16x2 integer SIMD (DPX).
//
// The literal form exists because ksw_extd2_sse widens the band to whole 16-lane SSE registers and lets the out-of-band lanes
// keep stale values that valid cells read once the band clips the matrix.  One GPU lane per SSE byte lane paid
// ~290 thread-instructions per true cell.  Here one lane owns TWO adjacent cells: every DP quantity is a pair of 16-bit halves
// in one register, adds / max / min are single VIADD.16x2 / VIMNMX(3).S16x2 / VIADDMNMX.S16x2 instructions, and the six state
// bytes of a cell pair travel as three 32-bit words (one LDS.128 / STS.128 per pair instead of 13 byte accesses per cell).
// Valid cells never leave the int8 range, but stale lanes do, so sums and differences are folded back to int8 where the
// reference's byte arithmetic would wrap (one byte permute each); see mmg_kswdpx_pair.
//
// Traceback direction (ksw2_extd2_sse.c:224-262): "first maximum wins" (left-aligned gaps) or "last maximum wins"
// (right-aligned) among z, a, b, a2, b2 -- one max over keys value*8 + preference gives value and direction at once.
#ifndef MMG_KSWDPX_H
#define MMG_KSWDPX_H
#include "mmg_core.h"

// ---- 16x2 primitives (device: one instruction each; host: the same arithmetic, for tests/emu)
MMG_HD uint32_t dpx_pack(int lo, int hi) { return (uint32_t)(lo & 0xffff) | (uint32_t)(hi & 0xffff) << 16; }
MMG_HD int dpx_lo(uint32_t a) { return (int16_t)(a & 0xffff); }
MMG_HD int dpx_hi(uint32_t a) { return (int16_t)(a >> 16); }
#ifdef __CUDA_ARCH__
MMG_HD uint32_t dpx_add(uint32_t a, uint32_t b) { return __vadd2(a, b); }
MMG_HD uint32_t dpx_sub(uint32_t a, uint32_t b) { return __vsub2(a, b); }
MMG_HD uint32_t dpx_max(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
MMG_HD uint32_t dpx_min(uint32_t a, uint32_t b) { return __vmins2(a, b); }
MMG_HD uint32_t dpx_max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
MMG_HD uint32_t dpx_addmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }   // max(a + b, c)
MMG_HD uint32_t dpx_minrelu(uint32_t a, uint32_t b) { return __vimin_s16x2_relu(a, b); }               // max(min(a, b), 0)
MMG_HD uint32_t dpx_prmt(uint32_t a, uint32_t b, uint32_t sel)
{ // not __byte_perm: that intrinsic masks the selector to 3 bits per byte and loses prmt's sign-replication bit
	uint32_t r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
	return r;
}
#else
#define DPX_EACH(expr_lo, expr_hi) dpx_pack((expr_lo), (expr_hi))
MMG_HD uint32_t dpx_add(uint32_t a, uint32_t b) { return DPX_EACH(dpx_lo(a) + dpx_lo(b), dpx_hi(a) + dpx_hi(b)); }
MMG_HD uint32_t dpx_sub(uint32_t a, uint32_t b) { return DPX_EACH(dpx_lo(a) - dpx_lo(b), dpx_hi(a) - dpx_hi(b)); }
MMG_HD int dpx_mx(int a, int b) { return a > b ? a : b; }
MMG_HD int dpx_mn(int a, int b) { return a < b ? a : b; }
MMG_HD uint32_t dpx_max(uint32_t a, uint32_t b) { return DPX_EACH(dpx_mx(dpx_lo(a), dpx_lo(b)), dpx_mx(dpx_hi(a), dpx_hi(b))); }
MMG_HD uint32_t dpx_min(uint32_t a, uint32_t b) { return DPX_EACH(dpx_mn(dpx_lo(a), dpx_lo(b)), dpx_mn(dpx_hi(a), dpx_hi(b))); }
MMG_HD uint32_t dpx_max3(uint32_t a, uint32_t b, uint32_t c) { return dpx_max(dpx_max(a, b), c); }
MMG_HD uint32_t dpx_addmax(uint32_t a, uint32_t b, uint32_t c) { return dpx_max(dpx_add(a, b), c); }
MMG_HD uint32_t dpx_minrelu(uint32_t a, uint32_t b) { return dpx_max(dpx_min(a, b), 0u); }
MMG_HD uint32_t dpx_prmt(uint32_t a, uint32_t b, uint32_t sel)
{ // PTX prmt, default mode: result byte i = source byte (sel nibble i & 7) of {b:a}; nibble bit 3 replicates that byte's sign
	const uint64_t src = (uint64_t)b << 32 | a;
	uint32_t r = 0;
	for (int i = 0; i < 4; ++i) {
		const uint32_t n = sel >> (4 * i) & 0xf;
		uint32_t byte = (uint32_t)(src >> (8 * (n & 7))) & 0xff;
		if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
		r |= byte << (8 * i);
	}
	return r;
}
#undef DPX_EACH
#endif

// state of a pair of cells t, t+1 (t even): three words of four int8 each
//   w0 = u[t] | u[t+1]<<8 | v[t]<<16 | v[t+1]<<24,  w1 = x | y (same order),  w2 = x2 | y2
struct KswPair { uint32_t w0, w1, w2; };
// byte offsets of a single cell's value inside the 16-byte slot of its pair (slot = t >> 1)
#define KSWDPX_U(t)  ((size_t)((t) >> 1) * 16 + 0 + ((t) & 1))
#define KSWDPX_V(t)  ((size_t)((t) >> 1) * 16 + 2 + ((t) & 1))
#define KSWDPX_X(t)  ((size_t)((t) >> 1) * 16 + 4 + ((t) & 1))
#define KSWDPX_Y(t)  ((size_t)((t) >> 1) * 16 + 6 + ((t) & 1))
#define KSWDPX_X2(t) ((size_t)((t) >> 1) * 16 + 8 + ((t) & 1))
#define KSWDPX_Y2(t) ((size_t)((t) >> 1) * 16 + 10 + ((t) & 1))

// per-job constants of the pair kernel, all as 16x2 pairs
struct KswDpxConst { uint32_t mch, q, q2, neg_qe, neg_qe2, pr[5]; };

template <int kMode>
MMG_HD KswDpxConst mmg_kswdpx_const(const KswGeom &g)
{
	KswDpxConst c;
	c.mch = dpx_pack(g.sc_mch, g.sc_mch), c.q = dpx_pack(g.q, g.q), c.q2 = dpx_pack(g.q2, g.q2);
	c.neg_qe = dpx_pack(-(g.q + g.e), -(g.q + g.e)), c.neg_qe2 = dpx_pack(-(g.q2 + g.e2), -(g.q2 + g.e2));
	// preference of candidate i (z, a, b, a2, b2) among equal values: left-aligned keeps the first (ksw2_extd2_sse.c:224-242,
	// strict '>'), right-aligned the last (:264-282, '>=' as "z > a ? d : 1")
	for (int i = 0; i < 5; ++i) { const int p = kMode == 2 ? i : 7 - i; c.pr[i] = dpx_pack(p, p); }
	return c;
}

// what a pair hands to the pair on its right: x[t+1] | v[t+1]<<8 | x2[t+1]<<16 of the OLD state
MMG_HD uint32_t mmg_kswdpx_carry(const KswPair &o) { return dpx_prmt(dpx_prmt(o.w1, o.w0, 0x0071), o.w2, 0x0510); }
MMG_HD uint32_t mmg_kswdpx_carry_of(int x, int v, int x2) { return (uint32_t)(x & 0xff) | (uint32_t)(v & 0xff) << 8 | (uint32_t)(x2 & 0xff) << 16; }

// two cells of ksw2_extd2_sse.c:184-313.  o: old state of the pair; s2: its two scores (bytes); prev: carry of the pair on the
// left (or the block's carry-in).  Returns the new state; *d2 = the two traceback bytes (kMode != 0).
template <int kMode, bool kFold = true /* false: the fast forms, where no stale lane is read and valid cells stay inside int8 */>
MMG_HD KswPair mmg_kswdpx_pair(const KswDpxConst &c, const KswPair &o, uint32_t s2, uint32_t prev, uint32_t *d2)
{
	// sign-extended operands
	const uint32_t xt1 = dpx_prmt(prev, o.w1, 0xC480), vt1 = dpx_prmt(prev, o.w0, 0xE691), x2t1 = dpx_prmt(prev, o.w2, 0xC4A2);
	const uint32_t ut = dpx_prmt(o.w0, 0, 0x9180), yt = dpx_prmt(o.w1, 0, 0xB3A2), y2t = dpx_prmt(o.w2, 0, 0xB3A2);
	uint32_t z = dpx_prmt(s2, 0, 0x9180);
	// The reference's lanes are int8 and the STALE ones do overflow (a lane left of the band keeps re-evaluating the recurrence on
	// its own old outputs and drifts: y = 94, 106, -150 -> 106 ...), so every sum and difference is folded back to int8 the way
	// _mm_add_epi8 / _mm_sub_epi8 wrap: one byte permute with sign replication per result.
#define W8(v) (kFold ? dpx_prmt((v), 0, 0xA280) : (v))
	uint32_t a = W8(dpx_add(xt1, vt1)), b = W8(dpx_add(yt, ut)), a2 = W8(dpx_add(x2t1, vt1)), b2 = W8(dpx_add(y2t, ut));
	uint32_t d = 0;
	if (kMode == 0) z = dpx_max3(dpx_max3(z, a, b), a2, b2);
	else { // value * 8 + preference: the largest key carries the value and which candidate it was
#define KSWDPX_KEY(v, i) (((v) << 3 & 0xfff8fff8u) | c.pr[i])
		const uint32_t k = dpx_max3(dpx_max3(KSWDPX_KEY(z, 0), KSWDPX_KEY(a, 1), KSWDPX_KEY(b, 2)), KSWDPX_KEY(a2, 3), KSWDPX_KEY(b2, 4));
#undef KSWDPX_KEY
		z = dpx_prmt(k >> 3, 0, 0xA280);                    // the int8 value back in both halves
		d = kMode == 2 ? (k & 0x00070007u) : ((k & 0x00070007u) ^ 0x00070007u);
	}
	z = dpx_min(z, c.mch);
	const uint32_t un = dpx_sub(z, vt1), vn = dpx_sub(z, ut);   // stored as bytes below: the pack keeps the low 8 bits
	uint32_t tmp = W8(dpx_sub(z, c.q));
	a = W8(dpx_sub(a, tmp)), b = W8(dpx_sub(b, tmp));
	tmp = W8(dpx_sub(z, c.q2));
	a2 = W8(dpx_sub(a2, tmp)), b2 = W8(dpx_sub(b2, tmp));
#undef W8
	// x = max(a, 0) - (q + e) = max(a - (q + e), -(q + e)); the gap continues iff a > 0 (left-aligned) or a >= 0 (right-aligned)
	const uint32_t xn = dpx_addmax(a, c.neg_qe, c.neg_qe), yn = dpx_addmax(b, c.neg_qe, c.neg_qe);
	const uint32_t x2n = dpx_addmax(a2, c.neg_qe2, c.neg_qe2), y2n = dpx_addmax(b2, c.neg_qe2, c.neg_qe2);
	if (kMode) {
		const uint32_t one = 0x00010001u;
		uint32_t fa, fb, fa2, fb2;
		if (kMode == 1) fa = dpx_minrelu(a, one), fb = dpx_minrelu(b, one), fa2 = dpx_minrelu(a2, one), fb2 = dpx_minrelu(b2, one);
		else fa = dpx_minrelu(dpx_add(a, one), one), fb = dpx_minrelu(dpx_add(b, one), one), fa2 = dpx_minrelu(dpx_add(a2, one), one), fb2 = dpx_minrelu(dpx_add(b2, one), one);
		d |= fa << 3 | fb << 4 | fa2 << 5 | fb2 << 6;
		*d2 = dpx_prmt(d, 0, 0x4420); // low byte of each half
	}
	KswPair n;
	n.w0 = dpx_prmt(un, vn, 0x6420), n.w1 = dpx_prmt(xn, yn, 0x6420), n.w2 = dpx_prmt(x2n, y2n, 0x6420);
	return n;
}

// ---- memory of one job in the pair layout: pairs (16 B per two cells) | s | sf | qr (contiguous, as the reference's s|sf|qr:
// the score refresh spills from s into sf and reads from sf into qr, ksw2_extd2_sse.c:158-172) | H
MMG_HD size_t mmg_kswdpx_lane_bytes(int qlen, int tlen) { return (size_t)((tlen + 15) / 16) * 16 * 10 + (size_t)((qlen + 15) / 16) * 16 + 16; }
MMG_HD size_t mmg_kswdpx_mem_bytes(int qlen, int tlen) { return ((mmg_kswdpx_lane_bytes(qlen, tlen) + 15) & ~(size_t)15) + (size_t)((tlen + 15) / 16) * 64; }

// The whole job walked by ONE thread in the pair layout, eight pairs per 16-lane block in lane order -- what the eight lanes of
// k_ksw_dpx do together.  The caller has filled sf[] (target, zero padded) and qr[] (reversed query, zero padded).
template <int kMode>
MMG_HDN inline void mmg_kswdpx_scalar_run(const KswGeom &g, int flag, int zdrop, uint8_t *mem, int32_t *H, uint8_t *p, KswEz &ez)
{
	const int tl16 = g.tlen_ * 16, qlen = g.qlen, tlen = g.tlen;
	uint32_t *PK = reinterpret_cast<uint32_t*>(mem);
	int8_t *sm = reinterpret_cast<int8_t*>(mem), *s = sm + (size_t)tl16 * 8;
	const uint8_t *sf = reinterpret_cast<const uint8_t*>(s + tl16), *qr = sf + tl16;
	const bool approx = (flag & MMG_EZ_APPROX_MAX) != 0;
	const int8_t n1 = (int8_t)(-g.q - g.e), n2 = (int8_t)(-g.q2 - g.e2);
	const KswDpxConst cst = mmg_kswdpx_const<kMode>(g);
	{
		const uint32_t w1 = 0x01010101u * (uint8_t)n1, w2 = 0x01010101u * (uint8_t)n2;
		for (int i = 0; i < tl16 / 2; ++i) PK[4 * i] = w1, PK[4 * i + 1] = w1, PK[4 * i + 2] = w2, PK[4 * i + 3] = 0;
		for (int i = 0; i < tl16; ++i) s[i] = 0, H[i] = MMG_KSW_NEG_INF;
	}
	int last_st = -1, last_en = -1;
	int32_t H0 = 0, last_H0_t = 0;
	for (int r = 0; r < qlen + tlen - 1; ++r) {
		int st0, en0;
		if (!mmg_ksw_band(g, r, &st0, &en0)) { ez.zdropped = 1; break; }
		const int st = st0 / 16 * 16, en = (en0 + 16) / 16 * 16 - 1;
		int x1, x21, v1;
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en) x1 = sm[KSWDPX_X(st - 1)], x21 = sm[KSWDPX_X2(st - 1)], v1 = sm[KSWDPX_V(st - 1)];
			else x1 = n1, x21 = n2, v1 = n1;
		} else x1 = n1, x21 = n2, v1 = mmg_ksw_first_col(g, r);
		if (en >= r) sm[KSWDPX_Y(r)] = n1, sm[KSWDPX_Y2(r)] = n2, sm[KSWDPX_U(r)] = (int8_t)mmg_ksw_first_col(g, r);
		{
			const uint8_t *qrr = qr + (qlen - 1 - r);
			for (int t = st0; t <= en0; t += 16)
				for (int l = 0; l < 16; ++l) s[t + l] = mmg_ksw_score(g, sf[t + l], qrr[t + l]);
		}
		uint32_t cin = mmg_kswdpx_carry_of(x1, v1, x21);
		for (int blk = st / 16; blk <= en / 16; ++blk) {
			uint32_t prev = cin;
			for (int l8 = 0; l8 < 8; ++l8) {
				const int pi = blk * 8 + l8;
				const KswPair o = {PK[4 * pi], PK[4 * pi + 1], PK[4 * pi + 2]};
				const uint32_t s2 = (uint32_t)(uint8_t)s[2 * pi] | (uint32_t)(uint8_t)s[2 * pi + 1] << 8;
				uint32_t d2 = 0;
				const KswPair n = mmg_kswdpx_pair<kMode>(cst, o, s2, prev, &d2);
				prev = mmg_kswdpx_carry(o);
				PK[4 * pi] = n.w0, PK[4 * pi + 1] = n.w1, PK[4 * pi + 2] = n.w2;
				if (kMode) { uint8_t *pr = p + ((size_t)r * g.n_col_ + (blk - st / 16)) * 16 + 2 * l8; pr[0] = (uint8_t)d2, pr[1] = (uint8_t)(d2 >> 8); }
			}
			cin = prev;
		}
		if (!approx) {
			int32_t max_H, max_t, H_en0;
			if (r > 0) {
				H_en0 = en0 > 0 ? H[en0 - 1] + sm[KSWDPX_U(en0)] : H[en0] + sm[KSWDPX_V(en0)];
				int32_t bh = H_en0, bt = en0; uint32_t br = 0;
				for (int t = st0; t < en0; ++t) {
					const int32_t h = H[t] + sm[KSWDPX_V(t)];
					H[t] = h;
					const uint32_t rk = mmg_ksw_max_rank(t, st0, en0);
					if (h > bh || (h == bh && rk < br)) bh = h, bt = t, br = rk;
				}
				H[en0] = H_en0;
				max_H = bh, max_t = bt;
			} else { H_en0 = sm[KSWDPX_V(0)] - g.qe_pre; H[0] = H_en0; max_H = H_en0, max_t = 0; }
			if (en0 == tlen - 1 && H_en0 > ez.mte) ez.mte = H_en0, ez.mte_q = r - en;
			if (r - st0 == qlen - 1) { const int32_t hs = H[st0]; if (hs > ez.mqe) ez.mqe = hs, ez.mqe_t = st0; }
			if (mmg_ksw_zdrop(&ez, max_H, r, max_t, zdrop, g.e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H[tlen - 1];
		} else {
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					const int32_t d0 = sm[KSWDPX_V(last_H0_t)], d1 = sm[KSWDPX_U(last_H0_t + 1)];
					if (d0 > d1) H0 += d0; else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) H0 += sm[KSWDPX_V(last_H0_t)];
				else ++last_H0_t, H0 += sm[KSWDPX_U(last_H0_t)];
			} else H0 = sm[KSWDPX_V(0)] - g.qe_pre, last_H0_t = 0;
			if ((flag & MMG_EZ_APPROX_DROP) && mmg_ksw_zdrop(&ez, H0, r, last_H0_t, zdrop, g.e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H0;
		}
		last_st = st, last_en = en;
	}
}

MMG_HDN inline void mmg_kswdpx_scalar(const KswGeom &g, int flag, int zdrop, int end_bonus, uint8_t *mem, int32_t *H, uint8_t *p, KswEz *ez_out, uint32_t *cigar)
{
	KswEz ez;
	mmg_ksw_reset(&ez);
	const bool with_cigar = !(flag & MMG_EZ_SCORE_ONLY);
	if (!with_cigar) mmg_kswdpx_scalar_run<0>(g, flag, zdrop, mem, H, p, ez);
	else if (!(flag & MMG_EZ_RIGHT)) mmg_kswdpx_scalar_run<1>(g, flag, zdrop, mem, H, p, ez);
	else mmg_kswdpx_scalar_run<2>(g, flag, zdrop, mem, H, p, ez);
	int i0, j0;
	if (with_cigar && mmg_ksw_trace_start(g, flag, end_bonus, &ez, &i0, &j0))
		ez.n_cigar = mmg_ksw_backtrack(g, !!(flag & MMG_EZ_REV_CIGAR), p, i0, j0, cigar);
	*ez_out = ez;
}
#endif
