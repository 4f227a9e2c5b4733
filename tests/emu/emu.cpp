// CPU emulation harness (TEST ONLY): compiles the __host__ __device__ building blocks of
// airlift_b200/csrc/mmg_core.h with g++ so the `-m "not gpu"` tests can check the very code the
// kernels execute against the oracle without a GPU.  Never linked into libmm2b200.so.
#include <vector>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <cstdio>
static long g_ksw_range_viol; // values of valid cells that left the int8 range (must stay 0: the fast form computes in 32 bits)
#define MMG_KSW_RANGE(v) do { if ((v) < -128 || (v) > 127) ++g_ksw_range_viol; } while (0)
#include "mmg_core.h"
#include "mmg_regheap.h"
#include "mmg_kswdpx.h"
#include "mmg_sketchwarp.h"
#include "mmg_rswarp.h"
#include "mmg_kswfast2.h"

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

// collect_seed_hits_heap the way the device does it for fragments whose heap order matters: plan -> expand (key = position) ->
// sort -> tie-group ranks -> mmg_heap_replay_ranks -> anchors in pop order, forward strand first (map.c:149-213)
int64_t emu_collect_ranked(void *idx, int64_t flag, int max_occ, int n_mv, const mm128 *mv, int qlen, mm128 *a, int64_t a_cap, int *rep_len, int *n_mini)
{
	EmuIdx *e = (EmuIdx*)idx;
	std::vector<int32_t> m_n(n_mv + 1), m_aoff(n_mv + 1); std::vector<uint64_t> m_val(n_mv + 1);
	for (int i = 0; i < n_mv; ++i) m_n[i] = mmg_idx_probe(e->v, mv[i].x >> 8, &m_val[i]);
	const int64_t n_a = mmg_frag_plan(mv, m_n.data(), n_mv, max_occ, rep_len, n_mini, nullptr);
	if (n_a > a_cap) return n_a;
	std::vector<int32_t> first, cnt, list_m;
	int32_t run = 0;
	for (int i = 0; i < n_mv; ++i) { m_aoff[i] = run; if (m_n[i] > 0 && m_n[i] < max_occ) { first.push_back(run), cnt.push_back(m_n[i]), list_m.push_back(i); run += m_n[i]; } }
	if ((int)first.size() > 256 || n_a >= (1 << 24)) return -1;
	std::vector<uint64_t> key(n_a); std::vector<int32_t> slot_m(n_a), slot_i(n_a), ord(n_a);
	for (size_t j = 0; j < first.size(); ++j)
		for (int i = 0; i < cnt[j]; ++i) { const int s = first[j] + i; key[s] = mmg_hit_pos(e->v.pos, m_n[list_m[j]], m_val[list_m[j]], (uint32_t)i); slot_m[s] = list_m[j], slot_i[s] = i; ord[s] = s; }
	std::sort(ord.begin(), ord.end(), [&](int x, int y) { return key[x] < key[y] || (key[x] == key[y] && x > y); }); // ties in any order: here, reversed
	std::vector<uint32_t> K(n_a + 1), heap(first.size() + 1), cur(first.size() + 1), pop(n_a + 1);
	for (int64_t g = 0; g < n_a; ++g) { int64_t l = g; while (l > 0 && key[ord[l - 1]] == key[ord[g]]) --l; K[ord[g]] = (uint32_t)l; }
	const int64_t np = mmg_heap_replay_ranks((int)first.size(), first.data(), cnt.data(), K.data(), heap.data(), cur.data(), pop.data());
	if (np != n_a) return -2;
	std::vector<mm128> fw, rv;
	for (int64_t t = 0; t < n_a; ++t) {
		const int m = slot_m[pop[t]];
		const uint64_t r = key[pop[t]];
		if (mmg_skip_seed(flag, r, (uint32_t)mv[m].y)) continue;
		const mm128 an = mmg_make_anchor(r, mv[m], mmg_is_tandem(mv, n_mv, m), qlen);
		((r & 1) == ((uint32_t)mv[m].y & 1) ? fw : rv).push_back(an);
	}
	std::copy(fw.begin(), fw.end(), a);
	std::copy(rv.begin(), rv.end(), a + fw.size());
	return (int64_t)(fw.size() + rv.size());
}

// K1, warp form (mmg_sketchwarp.h), lanes emulated one after the other; returns -1 when the read must take the state machine
int emu_sketch_warp(const char *str, int len, int w, int k, uint32_t rid, mm128 *out, int cap)
{
	std::vector<uint32_t> S((len + 7) / 8 + 4, 0x44444444u);
	for (int i = 0; i < len; ++i) { const int c = mmg_nt4((uint8_t)str[i]); S[i >> 3] = (S[i >> 3] & ~(0xfu << ((i & 7) * 4))) | (uint32_t)c << ((i & 7) * 4); }
	SketchUnit u; u.off = 0, u.len = len, u.rid = rid, u.y_add = 0, u.emit_start = 0, u.emit_end = len;
	if (!skw_eligible(u, w, k, 0)) return -1;
	std::vector<uint64_t> words(SKW_MAX_LEN / 32 + 2), xs(SKW_MAX_LEN);
	std::vector<uint8_t> zs(SKW_MAX_LEN);
	std::vector<mm128> tmp(len + 8);
	WarpEmu wp;
	const int n = mmg_sketch_warp<WarpEmu, true>(wp, S.data(), u, w, k, words.data(), xs.data(), zs.data(), tmp.data());
	if (n < 0) return -1;
	const int nc = mmg_sketch_warp<WarpEmu, false>(wp, S.data(), u, w, k, words.data(), xs.data(), zs.data(), nullptr);
	if (nc != n) return -3;
	if (n > cap) return -2;
	for (int i = 0; i < n; ++i) out[i] = tmp[i];
	return n;
}

// the heap merge on ranks: nreg == 0 runs the serial reference (mmg_heap_replay_ranks), nreg in {1, 2, 4} the warp form with the heap
// in registers (mmg_regheap.h), lanes emulated one after the other.  first/cnt: the lists' slots; K: rank of every slot.
int64_t emu_heap_replay(int n_lists, const int32_t *first, const int32_t *cnt, const uint32_t *K, uint32_t *pop, int nreg)
{
	std::vector<uint32_t> heap(n_lists + 1), cur(n_lists + 1);
	if (nreg == 0) return mmg_heap_replay_ranks(n_lists, first, cnt, K, heap.data(), cur.data(), pop);
	if (n_lists > 32 * nreg - 1) return -1;
	for (int j = 0; j < n_lists; ++j) heap[j] = K[first[j]] << 8 | (uint32_t)j, cur[j] = 0;
	if (n_lists > 1) for (int j = n_lists / 2 - 1; j >= 0; --j) mmg_rank_heap_down((uint32_t)j, (uint32_t)n_lists, heap.data());
	auto adv = [&](uint32_t j, int64_t t) -> uint32_t {
		pop[t] = (uint32_t)first[j] + cur[j];
		if (cur[j] + 1 < (uint32_t)cnt[j]) { ++cur[j]; return K[(uint32_t)first[j] + cur[j]]; }
		return RH_NONE;
	};
	WarpEmu wp;
	if (nreg == 1) return regheap_replay<WarpEmu, 1>(wp, n_lists, heap.data(), adv);
	if (nreg == 2) return regheap_replay<WarpEmu, 2>(wp, n_lists, heap.data(), adv);
	if (nreg == 4) return regheap_replay<WarpEmu, 4>(wp, n_lists, heap.data(), adv);
	return -1;
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

// the warp form of the same replay (mmg_rswarp.h), 32 emulated lanes
void emu_rs_sort_warp_128x(mm128 *a, int64_t n)
{
	std::vector<RsFrame> st(n / 65 + 4);
	std::vector<mm128> tmp(n + 1);
	std::vector<int32_t> dst(n + 1);
	std::vector<uint8_t> dig(n + 1);
	int32_t head[256], tail[256];
	WarpEmu wp;
	mmg_rs_sort_warp(wp, a, n, tmp.data(), dst.data(), dig.data(), st.data(), head, tail, KeyX());
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

// the literal form in the pair layout on 16x2 SIMD arithmetic (mmg_kswdpx.h), one thread walking the eight lanes of every block
int emu_ksw_dpx(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int q, int e, int q2, int e2, int w,
                int zdrop, int end_bonus, int flag, KswEz *ez_out, uint32_t *cigar)
{
	KswGeom g = mmg_ksw_geom(qlen, tlen, 5, mat, q, e, q2, e2, w);
	KswEz ez; mmg_ksw_reset(&ez);
	if (g.bail) { *ez_out = ez; return 0; }
	const int tl16 = g.tlen_ * 16;
	std::vector<uint32_t> mem32((mmg_kswdpx_mem_bytes(qlen, tlen) + 64) / 4, 0);
	uint8_t *mem = (uint8_t*)mem32.data();
	std::vector<int32_t> H(tl16);
	std::vector<uint8_t> p(((size_t)(qlen + tlen - 1) * g.n_col_ + 1) * 16);
	uint8_t *sf = mem + (size_t)tl16 * 9, *qr = sf + tl16;
	for (int i = 0; i < tlen; ++i) sf[i] = target[i];
	for (int i = 0; i < qlen; ++i) qr[i] = query[qlen - 1 - i];
	mmg_kswdpx_scalar(g, flag, zdrop, end_bonus, mem, H.data(), p.data(), ez_out, cigar);
	return 0;
}

// the pair form of the fast path (mmg_kswfast2.h): returns -1 when the job does not qualify
int emu_ksw_fast2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int q, int e, int q2, int e2, int w,
                  int zdrop, int end_bonus, int flag, int stride, KswEz *ez_out, uint32_t *cigar)
{
	KswGeom g = mmg_ksw_geom(qlen, tlen, 5, mat, q, e, q2, e2, w);
	KswEz ez; mmg_ksw_reset(&ez);
	if (g.bail) { *ez_out = ez; return 0; }
	for (int r = 0; r < qlen + tlen - 1; ++r) if (mmg_ksw_band_clips(g, r)) return -1;
	if (!mmg_ksw_fast_ok(g)) return -1;
	const int Wp = mmg_ksw_fast2_wp(qlen, tlen);
	std::vector<KswSlot> S((size_t)Wp * stride);
	memset(S.data(), 0x5a, S.size() * sizeof(KswSlot)); // garbage: a valid cell must never depend on an unwritten slot
	std::vector<uint8_t> tb((size_t)tlen * stride, 9), qb((size_t)(qlen + 2) * stride, 9);
	std::vector<uint32_t> p(mmg_ksw_fast2_p_bytes(qlen, tlen) / 4 + 4, 0xa5a5a5a5u);
	for (int i = 0; i < tlen; ++i) tb[(size_t)i * stride] = target[i];
	for (int i = 0; i < qlen; ++i) qb[(size_t)(i + 1) * stride] = query[i];
	mmg_ksw_fast2(g, flag, zdrop, end_bonus, S.data(), [&](int t) { return tb[(size_t)t * stride]; }, qb.data(), stride, p.data(), ez_out, cigar);
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

// ---- the post-chaining stages (mmg_post.h) for ONE fragment, every step run the way the kernels of mmg_post.cu run it,
// with the DP jobs answered by the scalar ksw walk above.  reads: 0..4 codes in mapping orientation.
#include <math.h>
#define HIT_PRIM_CACHE 2 // tiny on purpose: the third primary of a fragment already takes the path the device uses beyond its cache
#include "mmg_post.h"

extern "C" int emu_post_frag(const HitOpt *opt, int idx_k, uint32_t hash, int n_segs, const int32_t *qlens, const uint8_t *const *reads, const uint8_t *flip,
                             int n_ref, const uint8_t *const *refs, const uint32_t *ref_len, int n_u, const uint64_t *u_in, const mm128 *a_in, int rep_len,
                             int32_t *n_regs_out, HitRec *regs_out, int regs_cap, uint32_t *xw_out, int64_t xw_cap, int64_t *xw_used)
{
	PostShard sh;
	memset(&sh, 0, sizeof(sh));
	sh.opt = *opt, sh.nf = 1, sh.n_seq = n_segs, sh.idx_k = idx_k;
	// log tables: the host's own logf, exactly what the product builds at start-up
	std::vector<float> ld(1 << 16), li(1 << 16);
	for (int i = 0; i < (1 << 16); ++i) ld[i] = logf((float)i / opt->a), li[i] = logf((float)i);
	sh.lt.ld = ld.data(), sh.lt.li = li.data(), sh.lt.n = 1 << 16;
	// batch
	std::vector<int32_t> n_seg(1, n_segs), seg_off(1, 0), seq_len(qlens, qlens + n_segs);
	std::vector<uint64_t> q_off(n_segs + 1, 0);
	for (int j = 0; j < n_segs; ++j) q_off[j + 1] = q_off[j] + ((uint64_t)qlens[j] + 7) / 8 * 8;
	std::vector<uint32_t> Q(q_off[n_segs] / 8 + 4, 0);
	for (int j = 0; j < n_segs; ++j) for (int i = 0; i < qlens[j]; ++i) Q[(q_off[j] + i) >> 3] |= (uint32_t)reads[j][i] << (((q_off[j] + i) & 7) << 2);
	std::vector<uint64_t> ref_off(n_ref + 1, 0);
	for (int j = 0; j < n_ref; ++j) ref_off[j + 1] = ref_off[j] + ref_len[j];
	std::vector<uint32_t> S(ref_off[n_ref] / 8 + 4, 0);
	for (int j = 0; j < n_ref; ++j) for (uint32_t i = 0; i < ref_len[j]; ++i) S[(ref_off[j] + i) >> 3] |= (uint32_t)refs[j][i] << (((ref_off[j] + i) & 7) << 2);
	sh.n_seg = n_seg.data(), sh.seg_off = seg_off.data(), sh.seq_len = seq_len.data(), sh.q_off = q_off.data(), sh.flip = flip;
	sh.Q = Q.data(), sh.S = S.data(), sh.ref_off = ref_off.data(), sh.ref_len = ref_len;
	// chains
	int64_t n_v = 0;
	for (int i = 0; i < n_u; ++i) n_v += (int32_t)u_in[i];
	std::vector<int32_t> nu(1, n_u), rep(1, rep_len);
	std::vector<int64_t> uoff = {0, n_u}, voff = {0, n_v};
	std::vector<uint64_t> u(u_in, u_in + n_u);
	std::vector<mm128> a(a_in, a_in + n_v), a1(n_v + 1);
	std::vector<uint32_t> hv(1, hash);
	sh.nu = nu.data(), sh.rep = rep.data(), sh.uoff = uoff.data(), sh.voff = voff.data(), sh.u = u.data(), sh.a = a.data(), sh.hash = hv.data(), sh.a1 = a1.data();
	const int m = n_u + 1;
	std::vector<uint64_t> key_in(m), key(m), ascnt(m), cov(m);
	std::vector<HitRec> r0(m);
	std::vector<int32_t> w(m), n0(1);
	std::vector<mm128> big(m);
	std::vector<RsFrame> stack(m / 65 + 4);
	sh.key_in = key_in.data(), sh.key = key.data(), sh.ascnt = ascnt.data(), sh.r0 = r0.data(), sh.w = w.data(), sh.cov = cov.data(), sh.big = big.data(), sh.stack = stack.data(), sh.n0 = n0.data();
	std::vector<int32_t> cap(n_segs + 1, 0), n_reg(n_segs + 1, 0);
	std::vector<int64_t> roff(n_segs + 2, 0), a1_off(n_segs + 1, 0);
	std::vector<unsigned int> ctr(4, 0), n_jobs(1, 0);
	sh.cap = cap.data(), sh.n_reg = n_reg.data(), sh.a1_off = a1_off.data(), sh.ctr = ctr.data(), sh.n_jobs = n_jobs.data();
	// H1: keys as the key kernel writes them, the order the stable descending sort gives (emulated by its specification), records
	{
		std::vector<int64_t> pre(n_u + 1, 0);
		for (int i = 0; i < n_u; ++i) pre[i + 1] = pre[i] + (int32_t)u[i];
		std::vector<uint64_t> val_in(m);
		for (int g = 0; g < n_u; ++g) post_chain_key(sh, g, pre.data(), key_in.data(), val_in.data());
		std::vector<int> ord(n_u);
		for (int i = 0; i < n_u; ++i) ord[i] = i;
		std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return key_in[x] > key_in[y]; });
		for (int i = 0; i < n_u; ++i) key[i] = key_in[ord[i]], ascnt[i] = val_in[ord[i]];
		for (int g = 0; g < n_u; ++g) post_hit_record(sh, g);
	}
	if (getenv("EMU_WARP") && n_u >= atoi(getenv("EMU_WARP"))) { std::vector<int32_t> fast(HIT_FAST_WORDS + 2); post_hits_select_warp(WarpEmu(), sh, 0, fast.data()); } // the warp-cooperative form, lanes emulated one after the other
	else post_hits_select(sh, 0);
	for (int j = 0; j < n_segs; ++j) roff[j + 1] = roff[j] + cap[j];
	const int64_t slots = roff[n_segs] + 1;
	std::vector<HitRec> r1(slots), tmp1(slots);
	std::vector<RegionPlan> pl(slots);
	std::vector<uint32_t> xsize(slots, 0);
	std::vector<int64_t> xoff(slots + 1, 0);
	std::vector<uint64_t> skey(slots);
	std::vector<int32_t> sidx(slots);
	std::vector<mm128> sbig(slots);
	std::vector<RsFrame> sstack(slots / 65 + 2 * n_segs + 4);
	sh.roff = roff.data(), sh.r1 = r1.data(), sh.tmp1 = tmp1.data(), sh.pl = pl.data(), sh.xsize = xsize.data(), sh.xoff = xoff.data();
	sh.skey = skey.data(), sh.sidx = sidx.data(), sh.sbig = sbig.data(), sh.sstack = sstack.data();
	post_mates(sh, 0);
	std::vector<uint32_t> xw;
	if (opt->flag & HIT_F_CIGAR) {
		int8_t mat[25];
		for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) mat[i * 5 + j] = (int8_t)aln_mat(*opt, i, j);
		for (int round = 0; round < 64; ++round) {
			std::vector<DpJob> jobs(3 * slots + 4);
			sh.jobs = jobs.data(), sh.job_cap = (unsigned int)jobs.size(), n_jobs[0] = 0, ctr[0] = 0;
			for (int j = 0; j < n_segs; ++j) post_plan(sh, 0, j);
			// K4, scalar
			std::vector<DpRes> res(n_jobs[0] + 1);
			std::vector<uint32_t> cig;
			for (unsigned int k = 0; k < n_jobs[0]; ++k) {
				const DpJob &jb = jobs[k];
				const SeqView v = post_seq_view(sh, jb.seq_id);
				std::vector<uint8_t> q(jb.q_len), t(jb.t_len);
				for (int i = 0; i < jb.q_len; ++i) q[i] = (uint8_t)aln_q(v, jb.q_rev, jb.q_start + i);
				for (int i = 0; i < jb.t_len; ++i) t[i] = (uint8_t)aln_t(v, jb.rid, jb.t_start + i);
				if (jb.reversed) { std::reverse(q.begin(), q.end()); std::reverse(t.begin(), t.end()); }
				std::vector<uint32_t> cg(jb.q_len + jb.t_len + 4);
				KswEz ez;
				emu_ksw(jb.q_len, q.data(), jb.t_len, t.data(), mat, opt->q, opt->e, opt->q2, opt->e2, jb.w, jb.zdrop, jb.end_bonus, jb.flag, &ez, cg.data());
				res[k].ez = ez, res[k].cigar_off = cig.size();
				cig.insert(cig.end(), cg.begin(), cg.begin() + ez.n_cigar);
			}
			cig.push_back(0);
			sh.dp.res = res.data(), sh.dp.cig = cig.data();
			for (int j = 0; j < n_segs; ++j) post_size(sh, j);
			int64_t tot = 0;
			for (int64_t s = 0; s < slots; ++s) { xoff[s] = tot; tot += xsize[s]; }
			sh.x_base = (int64_t)xw.size();
			xw.resize(xw.size() + tot + 1);
			sh.xw = xw.data();
			for (int j = 0; j < n_segs; ++j) post_build(sh, 0, j);
			if (getenv("EMU_DEBUG")) { fprintf(stderr, "round %d: jobs %u new %u err %u;", round, n_jobs[0], ctr[0], ctr[1]); for (int j = 0; j < n_segs; ++j) { fprintf(stderr, " read %d cap %d:", j, cap[j]); for (int i = 0; i < n_reg[j]; ++i) fprintf(stderr, " [st %d cnt %d as %d split %u]", pl[roff[j] + i].state, r1[roff[j] + i].cnt, r1[roff[j] + i].as, HB_SPLIT(r1[roff[j] + i].bits)); } fprintf(stderr, "\n"); }
			xw.pop_back();
			if (ctr[0] == 0) break;
		}
		sh.xw = xw.data();
		for (int j = 0; j < n_segs; ++j) post_final(sh, j);
	}
	sh.xw = xw.data();
	post_finish(sh, 0);
	if (ctr[1]) return -(int)ctr[1];
	int64_t used = 0;
	for (int j = 0; j < n_segs; ++j) {
		n_regs_out[j] = n_reg[j];
		if (n_reg[j] > regs_cap) return -100;
		for (int i = 0; i < n_reg[j]; ++i) {
			HitRec h = r1[roff[j] + i];
			if (h.p) {
				const HitExtra *x = hit_ext(xw.data(), h.p);
				const int64_t wds = (int64_t)(sizeof(HitExtra) / 4) + x->n_cigar;
				if (used + wds > xw_cap) return -101;
				memcpy(xw_out + used, x, (size_t)wds * 4);
				h.p = (uint64_t)used + 1, used += wds;
			}
			regs_out[(size_t)j * regs_cap + i] = h;
		}
	}
	*xw_used = used;
	return 0;
}
