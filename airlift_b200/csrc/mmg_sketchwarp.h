// mmg_sketchwarp.h -- K1 for short reads, one warp per read: mm_sketch (sketch.c:77-143) with the positions of the read
// spread over the lanes instead of a state machine stepping through them.
//
// For a read without ambiguous bases and an odd k (no k-mer equals its own reverse complement, so sketch.c:108 never skips)
// every position pushes exactly one entry into the w-entry ring, and what the state machine emits at position i depends only on
// the hashes of positions i-w .. i:
//     P  = the ring's minimum before position i      = rightmost smallest hash among positions [i-w, i-1]
//     A  (i == w+k-2, sketch.c:117-122)   entries equal to P other than P, oldest first
//     B  (hash_i <= P, sketch.c:123-125)  P itself, once the window has been full for a position (i >= w+k-1)
//     C  (P sits in the slot position i overwrites, sketch.c:126-137)   P, then the entries of [i-w+1, i] that tie with the new
//                                          minimum M (rightmost smallest of [i-w+1, i]) other than M, oldest first
//     end (sketch.c:141-142)              the minimum of the last window
// So: lanes hash 32 consecutive k-mers from 2-bit packed words (forward and reverse-complement k-mer from one 64-bit window:
// the window read low-base-first IS the complemented reverse strand register, its 2-bit-group reversal the forward one), keep
// the hashes in shared memory, and every lane decides the events of its own position with one scan over the w-1 hashes before
// it; a warp prefix sum over the event counts places the output.  Reads with N, even k, HPC or more than SKW_MAX_LEN bases take
// the per-thread state machine (mmg_sketch_unit), which is the same function for every input.
#ifndef MMG_SKETCHWARP_H
#define MMG_SKETCHWARP_H
#include "mmg_warp.h"

#define SKW_MAX_LEN 512

MMG_HD uint64_t skw_rev2(uint64_t x)
{ // reverse the order of the 32 two-bit groups of x
#ifdef __CUDA_ARCH__
	x = __brevll(x);
#else
	x = (x >> 32) | (x << 32);
	x = (x & 0xffff0000ffff0000ULL) >> 16 | (x & 0x0000ffff0000ffffULL) << 16;
	x = (x & 0xff00ff00ff00ff00ULL) >> 8 | (x & 0x00ff00ff00ff00ffULL) << 8;
	x = (x & 0xf0f0f0f0f0f0f0f0ULL) >> 4 | (x & 0x0f0f0f0f0f0f0f0fULL) << 4;
	x = (x & 0xccccccccccccccccULL) >> 2 | (x & 0x3333333333333333ULL) << 2;
	x = (x & 0xaaaaaaaaaaaaaaaaULL) >> 1 | (x & 0x5555555555555555ULL) << 1;
#endif
	return (x & 0xaaaaaaaaaaaaaaaaULL) >> 1 | (x & 0x5555555555555555ULL) << 1; // bit reversal also swapped the two bits of a group
}

// true when the unit can take the warp form
MMG_HD bool skw_eligible(const SketchUnit &un, int w, int k, int is_hpc)
{
	return !is_hpc && (k & 1) && k <= 28 && w >= 1 && w <= 64 && un.emit_start == 0 && un.emit_end == un.len && un.len <= SKW_MAX_LEN && (un.off & 7) == 0;
}

// Scratch per warp: words[SKW_MAX_LEN / 32 + 2] (2-bit bases), xs[SKW_MAX_LEN] (hash << 8 | k), zs[SKW_MAX_LEN] (strand bit).
// Returns the number of minimizers, or -1 if the read holds an ambiguous base (nothing was written then).
template <class W, bool kWrite>
MMG_HDN inline int mmg_sketch_warp(const W &wp, const uint32_t *S, const SketchUnit &un, int w, int k, uint64_t *words, uint64_t *xs, uint8_t *zs, mm128 *out)
{
	const int len = un.len;
	const uint64_t mask = (1ULL << 2 * k) - 1;
	if (len < k) return 0;
	// ---- 2-bit packing, 32 bases per word; a lane takes one word (its four 4-bit source words) per round
	const int n_words = (len + 31) >> 5;
	unsigned bad = 0;
	for (int w0 = 0; w0 < n_words + 1; w0 += 32) {
		bad |= wp.ballot([&](int l) {
			const int wi = w0 + l;
			if (wi > n_words) return false;
			uint64_t acc = 0;
			bool amb = false;
			if (wi < n_words)
				for (int q = 0; q < 4; ++q) {
					const int b0 = wi * 32 + q * 8; // first base of this source word
					if (b0 >= len) break;
					uint32_t s4 = S[(un.off + (uint64_t)b0) >> 3];
					const int nb = len - b0 < 8 ? len - b0 : 8;
					if (nb < 8) s4 &= (1u << 4 * nb) - 1u; // padding nibbles hold 4
					if (s4 & 0xccccccccu) amb = true;       // a code above 3
					// low two bits of every nibble, squeezed together
					uint32_t t = s4 & 0x33333333u;
					t = (t | t >> 2) & 0x0f0f0f0fu;
					t = (t | t >> 4) & 0x00ff00ffu;
					t = (t | t >> 8) & 0x0000ffffu;
					acc |= (uint64_t)t << (16 * q);
				}
			words[wi] = acc;
			return amb;
		});
	}
	if (bad) return -1;
	wp.each([](int) {}); // the words are visible to every lane from here on
	// ---- hashes of all k-mers
	for (int i0 = 0; i0 < len; i0 += 32)
		wp.each([&](int l) {
			const int i = i0 + l;
			if (i >= len) return;
			if (i < k - 1) { xs[i] = MMG_NONE, zs[i] = 0; return; }
			const int st = i - k + 1, wd = st >> 5, sh = (st & 31) * 2;
			const uint64_t lo = words[wd], hi = words[wd + 1];
			const uint64_t E = (sh ? (lo >> sh | hi << (64 - sh)) : lo) & mask; // base st in the low bits ... base i on top
			const uint64_t k1 = E ^ mask;                                        // reverse strand register (sketch.c:106)
			const uint64_t k0 = skw_rev2(E) >> (64 - 2 * k);                     // forward strand register (sketch.c:105)
			const int z = k0 < k1 ? 0 : 1;
			xs[i] = mmg_hash64(z ? k1 : k0, mask) << 8 | (uint64_t)k;
			zs[i] = (uint8_t)z;
		});
	// ---- events
	auto X = [&](int j) -> uint64_t { return j >= k - 1 ? xs[j] : MMG_NONE; };
	auto put = [&](int at, int j) {
		if (kWrite) { out[at].x = xs[j]; out[at].y = ((uint64_t)un.rid << 32 | (uint32_t)j << 1 | zs[j]) + un.y_add; }
	};
	int n = 0;
	const int first = w + k - 2; // position at which the ring is full of valid entries for the first time
	for (int i0 = first; i0 < len; i0 += 32) {
		typename W::template Var<int> cnt, pi, mi_, na, nc; // per lane: events, P, M, ties at A, ties at C
		wp.each([&](int l) {
			const int i = i0 + l;
			cnt(l) = 0, pi(l) = -1, mi_(l) = -1, na(l) = 0, nc(l) = 0;
			if (i >= len) return;
			// Q = rightmost smallest of [i-w+1, i-1], with its multiplicity
			uint64_t qx = MMG_NONE; int qi = -1, qc = 0;
			for (int j = i - w + 1; j < i; ++j) {
				const uint64_t x = X(j);
				if (x == MMG_NONE) continue;
				if (x < qx) qx = x, qi = j, qc = 1;
				else if (x == qx) qi = j, ++qc;
			}
			const uint64_t xo = X(i - w); // the entry position i overwrites
			uint64_t px = qx; int p = qi, pc = qc; // P over [i-w, i-1]
			if (xo != MMG_NONE) {
				if (xo < qx) px = xo, p = i - w, pc = 1;
				else if (xo == qx) ++pc;
			}
			const uint64_t xi = xs[i];
			int c = 0;
			if (i == first && p >= 0) na(l) = pc - 1, c += pc - 1;
			if (p < 0 || xi <= px) { // B
				if (i >= w + k - 1 && p >= 0) ++c, pi(l) = p;
			} else if (p == i - w) { // C
				++c, pi(l) = p;
				int m, mc; // M over [i-w+1, i]
				if (qi < 0 || xi < qx) m = i, mc = 1;
				else if (xi == qx) m = i, mc = qc + 1;
				else m = qi, mc = qc;
				mi_(l) = m, nc(l) = mc - 1, c += mc - 1;
			}
			cnt(l) = c;
		});
		// exclusive prefix of the event counts over the lanes
		typename W::template Var<int> inc;
		wp.regs([&](int l) { inc(l) = cnt(l); });
		for (int d = 1; d < 32; d <<= 1) {
			auto o = wp.shfl(inc, [d](int l) { return l - d; });
			wp.regs([&](int l) { if (l >= d) inc(l) += o(l); });
		}
		const int tot = wp.get(inc, 31);
		if (kWrite)
			wp.each([&](int l) {
				const int i = i0 + l;
				if (i >= len || cnt(l) == 0) return;
				int at = n + inc(l) - cnt(l);
				if (i == first && na(l) > 0) { // A: ties of P, oldest first; P is the rightmost smallest of [i-w, i-1] (recomputed: rare path)
					uint64_t px = MMG_NONE; int p = -1;
					for (int j = i - w; j < i; ++j) { const uint64_t x = X(j); if (x != MMG_NONE && x <= px) px = x, p = j; }
					for (int j = i - w + 1; j < i; ++j) if (X(j) == px && j != p) put(at++, j);
				}
				if (pi(l) >= 0) put(at++, pi(l));
				if (nc(l) > 0) { const int m = mi_(l); const uint64_t mx = xs[m]; for (int j = i - w + 1; j <= i; ++j) if (X(j) == mx && j != m) put(at++, j); }
			});
		n += tot;
	}
	{ // the minimum of the last window (sketch.c:141-142)
		uint64_t bx = MMG_NONE; int bj = -1;
		const int lo = len - w > k - 1 ? len - w : k - 1;
		for (int j = lo; j < len; ++j) { const uint64_t x = xs[j]; if (x <= bx) bx = x, bj = j; }
		if (bj >= 0) { wp.one([&]() { put(n, bj); }); ++n; }
	}
	return n;
}
#endif
