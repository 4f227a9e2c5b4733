// CPU emulation harness (TEST ONLY): compiles the __host__ __device__ building blocks of
// airlift_b200/csrc/mmg_core.h with g++ so the `-m "not gpu"` tests can check the very code the
// kernels execute against the oracle without a GPU.  Never linked into libmm2b200.so.
#include <vector>
#include <cstring>
#include <cstdlib>
#include "mmg_core.h"

extern "C" {

int emu_sketch(const uint8_t *ascii, int len, int w, int k, uint32_t rid, int is_hpc, int chunk, mm128 *out, int cap)
{
	std::vector<uint32_t> S((len + 7) / 8 + 4, 0);
	const uint64_t base = 5; // deliberately unaligned start inside the packed array
	S.resize((len + base + 7) / 8 + 4, 0);
	for (int i = 0; i < len; ++i) S[(base + i) >> 3] |= (uint32_t)mmg_nt4(ascii[i]) << (((base + i) & 7) << 2);
	int n = 0;
	if (is_hpc || chunk <= 0) chunk = len;
	std::vector<mm128> tmp(len + 8);
	for (int s = 0; s < len; s += chunk) {
		SketchUnit u; u.off = base, u.len = len, u.rid = rid, u.y_add = 0, u.emit_start = s, u.emit_end = s + chunk < len ? s + chunk : len;
		int c0 = mmg_sketch_unit<false>(S.data(), u, w, k, is_hpc, nullptr);
		int c1 = mmg_sketch_unit<true>(S.data(), u, w, k, is_hpc, tmp.data());
		if (c0 != c1) return -1;
		for (int i = 0; i < c1; ++i) { if (n < cap) out[n] = tmp[i]; ++n; }
	}
	return n;
}

struct EmuIdx { std::vector<IdxSlot> slots; std::vector<uint64_t> pos; IdxView v; };

// keys[] ascending (one entry per occurrence), pos[] the matching positions (ascending within a key)
void *emu_idx_new(int64_t n, const uint64_t *keys, const uint64_t *pos)
{
	EmuIdx *e = new EmuIdx();
	int64_t n_keys = 0;
	for (int64_t i = 0; i < n; ++i) if (i == 0 || keys[i] != keys[i - 1]) ++n_keys;
	uint64_t n_slots = 1024; int bits = 10;
	while (n_slots < (uint64_t)n_keys * 2) n_slots <<= 1, ++bits;
	e->slots.assign(n_slots, IdxSlot{MMG_SLOT_EMPTY, MMG_SLOT_EMPTY});
	e->pos.assign(pos, pos + n);
	const int shift = 64 - bits;
	for (int64_t i = 0; i < n;) {
		int64_t j = i; while (j < n && keys[j] == keys[i]) ++j;
		const uint32_t cnt = (uint32_t)(j - i);
		uint64_t s = mmg_slot_hash(keys[i], shift);
		while (e->slots[s].key != MMG_SLOT_EMPTY) s = (s + 1) & (n_slots - 1);
		e->slots[s].key = cnt == 1 ? (keys[i] | MMG_SLOT_SINGLE) : keys[i];
		e->slots[s].val = cnt == 1 ? pos[i] : ((uint64_t)i << 32 | cnt);
		i = j;
	}
	memset(&e->v, 0, sizeof(e->v));
	e->v.slots = e->slots.data(), e->v.pos = e->pos.data(), e->v.slot_mask = n_slots - 1, e->v.slot_shift = shift;
	return e;
}
void emu_idx_free(void *p) { delete (EmuIdx*)p; }

int64_t emu_collect(void *idx, int heap_sort, int64_t flag, int max_occ, int n_mv, const mm128 *mv, int qlen, mm128 *a, int64_t a_cap,
                    int *rep_len, int *n_mini, uint64_t *mini)
{
	EmuIdx *e = (EmuIdx*)idx;
	std::vector<int32_t> m_n(n_mv + 1); std::vector<uint64_t> m_val(n_mv + 1);
	for (int i = 0; i < n_mv; ++i) m_n[i] = mmg_idx_probe(e->v, mv[i].x >> 8, &m_val[i]);
	int64_t n_a = mmg_frag_plan(mv, m_n.data(), n_mv, max_occ, rep_len, n_mini, mini);
	if (n_a > a_cap) return n_a;
	std::vector<mm128> heap(n_mv + 1); std::vector<RsFrame> stack(n_a / 65 + 4);
	if (heap_sort) return mmg_fill_heap(mv, m_n.data(), m_val.data(), n_mv, max_occ, e->v.pos, flag, qlen, n_a, heap.data(), a);
	return mmg_fill_flat(mv, m_n.data(), m_val.data(), n_mv, max_occ, e->v.pos, flag, qlen, a, stack.data());
}

int emu_chain(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc, int is_cdna, int n_segs,
              int64_t n, mm128 *a, uint64_t *u_out)
{
	if (n <= 0) return 0;
	ChainParams P; P.max_dist_x = max_dist_x, P.max_dist_y = max_dist_y, P.bw = bw, P.max_skip = max_skip, P.max_iter = max_iter;
	P.min_cnt = min_cnt, P.min_sc = min_sc, P.is_cdna = is_cdna, P.n_segs = n_segs;
	std::vector<int32_t> work(4 * n); std::vector<uint64_t> u(2 * n); std::vector<mm128> b(n); std::vector<RsFrame> st(n / 65 + 4);
	int64_t n_v = 0;
	mmg_chain_fill_seq(P, n, a, work.data(), work.data() + n, work.data() + 2 * n, work.data() + 3 * n);
	int n_u = mmg_chain_backtrack(P, n, a, work.data(), work.data() + n, work.data() + 2 * n, work.data() + 3 * n, u.data(), b.data(), st.data(), &n_v);
	for (int i = 0; i < n_u; ++i) u_out[i] = u[i];
	return n_u;
}

void emu_rs_sort_128x(mm128 *a, int64_t n)
{
	std::vector<RsFrame> st(n / 65 + 4);
	mmg_rs_sort_exact(a, n, st.data(), KeyX());
}

// scalar walk over the DP using the device cell/band/rank/backtrack helpers (lanes replaced by loops)
int emu_ksw(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int q, int e, int q2, int e2, int w,
            int zdrop, int end_bonus, int flag, KswEz *ez_out, uint32_t *cigar)
{
	KswGeom g = mmg_ksw_geom(qlen, tlen, 5, mat, q, e, q2, e2, w);
	KswEz ez; mmg_ksw_reset(&ez);
	if (g.bail) { *ez_out = ez; return 0; }
	const int tl16 = g.tlen_ * 16;
	std::vector<int8_t> mem(mmg_ksw_mem_bytes(qlen, tlen) + 64, 0);
	std::vector<int32_t> H(tl16, MMG_KSW_NEG_INF);
	std::vector<uint8_t> p(((size_t)(qlen + tlen - 1) * g.n_col_ + 1) * 16);
	int8_t *u = mem.data(), *v = u + tl16, *x = v + tl16, *y = x + tl16, *x2 = y + tl16, *y2 = x2 + tl16, *s = y2 + tl16;
	uint8_t *sf = (uint8_t*)(s + tl16), *qr = sf + tl16;
	for (int i = 0; i < tl16; ++i) u[i] = v[i] = x[i] = y[i] = (int8_t)(-g.q - g.e), x2[i] = y2[i] = (int8_t)(-g.q2 - g.e2);
	for (int i = 0; i < tlen; ++i) sf[i] = target[i];
	for (int i = 0; i < qlen; ++i) qr[i] = query[qlen - 1 - i];
	const bool approx = flag & MMG_EZ_APPROX_MAX, with_cigar = !(flag & MMG_EZ_SCORE_ONLY);
	const int mode = !with_cigar ? 0 : !(flag & MMG_EZ_RIGHT) ? 1 : 2;
	int last_st = -1, last_en = -1; int32_t H0 = 0, last_H0_t = 0;
	for (int r = 0; r < qlen + tlen - 1; ++r) {
		int st0, en0;
		if (!mmg_ksw_band(g, r, &st0, &en0)) { ez.zdropped = 1; break; }
		const int st = st0 / 16 * 16, en = (en0 + 16) / 16 * 16 - 1;
		int x1, x21, v1;
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en) x1 = x[st - 1], x21 = x2[st - 1], v1 = v[st - 1];
			else x1 = -g.q - g.e, x21 = -g.q2 - g.e2, v1 = -g.q - g.e;
		} else x1 = -g.q - g.e, x21 = -g.q2 - g.e2, v1 = mmg_ksw_first_col(g, r);
		if (en >= r) y[r] = (int8_t)(-g.q - g.e), y2[r] = (int8_t)(-g.q2 - g.e2), u[r] = (int8_t)mmg_ksw_first_col(g, r);
		const uint8_t *qrr = qr + (qlen - 1 - r);
		for (int t = st0; t <= en0; t += 16) { int8_t tmp[16]; for (int l = 0; l < 16; ++l) tmp[l] = mmg_ksw_score(g, sf[t + l], qrr[t + l]); memcpy(s + t, tmp, 16); }
		for (int blk = st / 16; blk <= en / 16; ++blk) {
			int8_t xo[16], vo[16], x2o[16];
			for (int l = 0; l < 16; ++l) xo[l] = x[blk * 16 + l], vo[l] = v[blk * 16 + l], x2o[l] = x2[blk * 16 + l];
			for (int l = 0; l < 16; ++l) {
				const int i = blk * 16 + l;
				const int8_t xt1 = l ? xo[l - 1] : (int8_t)x1, vt1 = l ? vo[l - 1] : (int8_t)v1, x2t1 = l ? x2o[l - 1] : (int8_t)x21;
				KswCell c = mode == 0 ? mmg_ksw_cell<0>(g, s[i], xt1, vt1, u[i], y[i], x2t1, y2[i])
				          : mode == 1 ? mmg_ksw_cell<1>(g, s[i], xt1, vt1, u[i], y[i], x2t1, y2[i])
				                      : mmg_ksw_cell<2>(g, s[i], xt1, vt1, u[i], y[i], x2t1, y2[i]);
				u[i] = c.u, v[i] = c.v, x[i] = c.x, y[i] = c.y, x2[i] = c.x2, y2[i] = c.y2;
				if (mode) p[((size_t)r * g.n_col_ + (blk - st / 16)) * 16 + l] = c.d;
			}
			x1 = xo[15], v1 = vo[15], x21 = x2o[15];
		}
		if (!approx) {
			int32_t max_H, max_t, H_en0;
			if (r > 0) {
				H_en0 = en0 > 0 ? H[en0 - 1] + u[en0] : H[en0] + v[en0];
				int32_t bh = H_en0, bt = en0; uint32_t br = 0;
				for (int t = st0; t < en0; ++t) {
					const int32_t h = H[t] + v[t]; H[t] = h;
					const uint32_t rk = mmg_ksw_max_rank(t, st0, en0);
					if (h > bh || (h == bh && rk < br)) bh = h, bt = t, br = rk;
				}
				H[en0] = H_en0; max_H = bh, max_t = bt;
			} else { H_en0 = v[0] - g.qe_pre; H[0] = H_en0; max_H = H_en0, max_t = 0; }
			if (en0 == tlen - 1 && H_en0 > ez.mte) ez.mte = H_en0, ez.mte_q = r - en;
			if (r - st0 == qlen - 1) { const int32_t hs = H[st0]; if (hs > ez.mqe) ez.mqe = hs, ez.mqe_t = st0; }
			if (mmg_ksw_zdrop(&ez, max_H, r, max_t, zdrop, g.e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H[tlen - 1];
		} else {
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					const int32_t d0 = v[last_H0_t], d1 = u[last_H0_t + 1];
					if (d0 > d1) H0 += d0; else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) H0 += v[last_H0_t];
				else ++last_H0_t, H0 += u[last_H0_t];
			} else H0 = v[0] - g.qe_pre, last_H0_t = 0;
			if ((flag & MMG_EZ_APPROX_DROP) && mmg_ksw_zdrop(&ez, H0, r, last_H0_t, zdrop, g.e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H0;
		}
		last_st = st, last_en = en;
	}
	int i0, j0;
	if (with_cigar && mmg_ksw_trace_start(g, flag, end_bonus, &ez, &i0, &j0))
		ez.n_cigar = mmg_ksw_backtrack(g, !!(flag & MMG_EZ_REV_CIGAR), p.data(), i0, j0, cigar);
	*ez_out = ez;
	return 0;
}

}
