// mmg_regheap.h -- klib's binary heap (ksort.h:43-53) replayed by a whole warp, the heap in registers.
//
// Why: collect_seed_hits_heap (map.c:149-213) merges the position lists of a fragment's minimizers through a binary heap, and
// where two lists carry equal positions the order they leave the heap in is a property of that heap's history.
// A repeat-family fragment makes ~1.5 x 10^5 pops, each depending on the one before, so the time of a pop IS the time of the
// kernel.  One lane walking the heap in shared memory spends ~190 dependent instructions per pop (~0.45 us).  Here the heap's
// nodes live in the registers of the 32 lanes and one pop is a constant number of warp-wide steps, whatever the depth:
//   * every internal node knows which child is the smaller one (right only if strictly smaller, ksort.h:47); one ballot
//     publishes those bits, and a node is on the sift path of the root iff its ancestors' bits spell its own index -- two
//     logic ops against masks fixed at start-up;
//   * the sift stops at the first path node whose key is larger than the incoming key (ksort.h:48): a ballot and a find-first;
//   * the path nodes above the stop take their path child's key (already fetched for the direction bits), the last one the
//     incoming key.
// Nodes are numbered from 1 (children 2i, 2i+1, always in the same register), node i in lane i & 31 of register i >> 5;
// absent nodes hold RH_INF, which is larger than any key and so never moves.  Keys are rank << 8 | list (mmg_heap_replay_ranks).
#ifndef MMG_REGHEAP_H
#define MMG_REGHEAP_H
#include "mmg_warp.h"

#define RH_INF  0xffffffffu
#define RH_NONE 0xffffffffu   // advance(): the list has no further element

template <class W, int NREG> struct RegHeap {
	static constexpr int NI = NREG > 1 ? NREG / 2 : 1;  // registers that hold internal nodes = words of direction bits
	typename W::template Var<uint32_t> k[NREG], care[NREG][NI], need[NREG][NI];
};

// heap0: n keys already in heap order (0-based, as ks_heapmake leaves them)
template <class W, int NREG>
MMG_HDN inline void regheap_init(const W &wp, RegHeap<W, NREG> &h, const uint32_t *heap0, int n)
{
	constexpr int NI = RegHeap<W, NREG>::NI;
	wp.each([&](int l) {
#pragma unroll
		for (int r = 0; r < NREG; ++r) {
			const int i = l + 32 * r;
			h.k[r](l) = (i >= 1 && i <= n) ? heap0[i - 1] : RH_INF;
			uint32_t c[NI], d[NI];
#pragma unroll
			for (int w = 0; w < NI; ++w) c[w] = d[w] = 0;
			for (int a = i; a > 1; a >>= 1) { // ancestor a >> 1 must point at a
				const int p = a >> 1;
#pragma unroll
				for (int w = 0; w < NI; ++w)
					if (w == p >> 5) { c[w] |= 1u << (p & 31); if (a & 1) d[w] |= 1u << (p & 31); }
			}
#pragma unroll
			for (int w = 0; w < NI; ++w) h.care[r][w](l) = c[w], h.need[r][w](l) = d[w];
		}
	});
}

// ks_heapdown(0, n, l) after l[0] = tmp; returns the new root
template <class W, int NREG>
MMG_HDN inline uint32_t regheap_replace_root(const W &wp, RegHeap<W, NREG> &h, uint32_t tmp)
{
	constexpr int NI = RegHeap<W, NREG>::NI;
	typename W::template Var<uint32_t> lc[NI], rc[NI], dir[NI];
	uint32_t dmask[NI];
#pragma unroll
	for (int r = 0; r < NI; ++r) {
		// children of node l + 32 r: nodes 2l + 64r and 2l + 64r + 1, in register 2r + (l >> 4), lanes 2l & 31 and (2l + 1) & 31
		auto a0 = wp.shfl(h.k[NREG > 1 ? 2 * r : 0], [](int l) { return (2 * l) & 31; });
		auto b0 = wp.shfl(h.k[NREG > 1 ? 2 * r : 0], [](int l) { return (2 * l + 1) & 31; });
		if (NREG > 1) {
			auto a1 = wp.shfl(h.k[NREG > 1 ? 2 * r + 1 : 0], [](int l) { return (2 * l) & 31; });
			auto b1 = wp.shfl(h.k[NREG > 1 ? 2 * r + 1 : 0], [](int l) { return (2 * l + 1) & 31; });
			wp.regs([&](int l) { lc[r](l) = l < 16 ? a0(l) : a1(l), rc[r](l) = l < 16 ? b0(l) : b1(l); });
		} else wp.regs([&](int l) { lc[r](l) = l < 16 ? a0(l) : RH_INF, rc[r](l) = l < 16 ? b0(l) : RH_INF; });
		wp.regs([&](int l) { dir[r](l) = lc[r](l) > (rc[r](l) | 0xffu) ? 1u : 0u; }); // rank(left) > rank(right): the right child is strictly smaller
		dmask[r] = wp.ballot([&](int l) { return dir[r](l) != 0; });
	}
	const uint32_t tx = tmp | 0xffu; // key > tx  <=>  rank(key) > rank(tmp)
	typename W::template Var<uint32_t> on[NREG];
	uint32_t s = 32u * NREG; // first path node that stays where it is
#pragma unroll
	for (int r = NREG - 1; r >= 0; --r) {
		wp.regs([&](int l) {
			uint32_t miss = (r == 0 && l == 0) ? 1u : 0u;
#pragma unroll
			for (int w = 0; w < NI; ++w) miss |= (dmask[w] ^ h.need[r][w](l)) & h.care[r][w](l);
			on[r](l) = miss == 0 ? 1u : 0u;
		});
		const unsigned m = wp.ballot([&](int l) { return on[r](l) && !(r == 0 && l == 1) && h.k[r](l) > tx; });
		if (m) s = 32u * r + (uint32_t)mmg_ffs(m) - 1u;
	}
	wp.regs([&](int l) {
#pragma unroll
		for (int r = 0; r < NREG; ++r) {
			const uint32_t i = (uint32_t)l + 32u * r;
			if (on[r](l) && i < s) {
				uint32_t nk = tmp;
				if (r < NI && (NREG > 1 || l < 16)) { const uint32_t c = 2u * i + dir[r](l); if (c < s) nk = dir[r](l) ? rc[r](l) : lc[r](l); }
				h.k[r](l) = nk;
			}
		}
	});
	return wp.get(h.k[0], 1);
}

// The merge itself (mmg_heap_replay_ranks): advance(j, t) is called once per pop by the whole warp with the list j whose element
// leaves the heap as pop t; it records the pop and returns the rank of that list's next element, or RH_NONE.
template <class W, int NREG, class Adv>
MMG_HDN inline int64_t regheap_replay(const W &wp, int n_lists, const uint32_t *heap0, Adv &&advance)
{
	RegHeap<W, NREG> h;
	regheap_init(wp, h, heap0, n_lists);
	int hs = n_lists;
	int64_t t = 0;
	uint32_t root = hs ? heap0[0] : 0;
	while (hs > 0) {
		const uint32_t j = root & 0xffu;
		const uint32_t nk = advance(j, t++);
		uint32_t tmp;
		if (nk != RH_NONE) tmp = nk << 8 | j;
		else { // the last node takes the root's place (map.c:201-204)
			tmp = 0;
#pragma unroll
			for (int r = 0; r < NREG; ++r) {
				const uint32_t v = wp.get(h.k[r], hs & 31);
				if (r == hs >> 5) tmp = v;
			}
			wp.regs([&](int l) {
#pragma unroll
				for (int r = 0; r < NREG; ++r) if (r == hs >> 5 && l == (hs & 31)) h.k[r](l) = RH_INF;
			});
			if (--hs == 0) break;
		}
		root = regheap_replace_root(wp, h, tmp);
	}
	return t;
}
#endif
