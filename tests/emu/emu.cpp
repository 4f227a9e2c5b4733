// CPU emulation harness (TEST ONLY): compiles the __host__ __device__ building blocks of
// airlift_b200/csrc/mmg_core.h with g++ so the `-m "not gpu"` tests can check the very code the
// kernels execute against the oracle without a GPU.  Never linked into libmm2b200.so.
#include <vector>
#include <cstring>
#include <cstdlib>
static long g_ksw_range_viol; // values of valid cells that left the int8 range (must stay 0: the fast form computes in 32 bits)
#define MMG_KSW_RANGE(v) do { if ((v) < -128 || (v) > 127) ++g_ksw_range_viol; } while (0)
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

// the one-thread-per-job ksw walk the device runs for small jobs (mmg_ksw_scalar), on the CPU
int emu_ksw(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int q, int e, int q2, int e2, int w,
            int zdrop, int end_bonus, int flag, KswEz *ez_out, uint32_t *cigar)
{
	KswGeom g = mmg_ksw_geom(qlen, tlen, 5, mat, q, e, q2, e2, w);
	KswEz ez; mmg_ksw_reset(&ez);
	if (g.bail) { *ez_out = ez; return 0; }
	const int tl16 = g.tlen_ * 16;
	std::vector<int8_t> mem(mmg_ksw_mem_bytes(qlen, tlen) + 64, 0);
	std::vector<int32_t> H(tl16);
	std::vector<uint8_t> p(((size_t)(qlen + tlen - 1) * g.n_col_ + 1) * 16);
	uint8_t *sf = (uint8_t*)mem.data() + 7 * tl16, *qr = sf + tl16;
	for (int i = 0; i < tlen; ++i) sf[i] = target[i];
	for (int i = 0; i < qlen; ++i) qr[i] = query[qlen - 1 - i];
	mmg_ksw_scalar(g, flag, zdrop, end_bonus, mem.data(), H.data(), p.data(), ez_out, cigar);
	return 0;
}

// the fast form (band never clips): returns -1 when the job does not qualify
int emu_ksw_fast(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int q, int e, int q2, int e2, int w,
                 int zdrop, int end_bonus, int flag, int stride, KswEz *ez_out, uint32_t *cigar)
{
	KswGeom g = mmg_ksw_geom(qlen, tlen, 5, mat, q, e, q2, e2, w);
	KswEz ez; mmg_ksw_reset(&ez);
	if (g.bail) { *ez_out = ez; return 0; }
	for (int r = 0; r < qlen + tlen - 1; ++r) if (mmg_ksw_band_clips(g, r)) return -1;
	if (!mmg_ksw_fast_ok(g)) return -1;
	const int W = (qlen < tlen ? qlen : tlen) + 1;
	std::vector<uint64_t> S((size_t)W * stride, 0x5a5a5a5a5a5a5a5aULL); // garbage: the fast form must never read an unwritten slot
	std::vector<uint8_t> tb((size_t)tlen * stride, 9), qb((size_t)qlen * stride, 9);
	std::vector<uint32_t> p(mmg_ksw_fast_p_bytes(qlen, tlen) / 4 + 4, 0xa5a5a5a5u);
	for (int i = 0; i < tlen; ++i) tb[(size_t)i * stride] = target[i];
	for (int i = 0; i < qlen; ++i) qb[(size_t)i * stride] = query[i];
	const long viol0 = g_ksw_range_viol;
	mmg_ksw_fast(g, flag, zdrop, end_bonus, S.data(), tb.data(), qb.data(), stride, p.data(), ez_out, cigar);
	return g_ksw_range_viol != viol0 ? -2 : 0;
}

}
