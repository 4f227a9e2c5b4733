// mmg_rswarp.h -- klib's in-place MSD radix sort (ksort.h:101-151) replayed exactly by a WARP.
//
// Why a replay at all: radix_sort_128x is not stable, so where two anchors share a key the order they end in is a property of
// the in-place permutation; mmg_rs_sort_exact (mmg_core.h) replays it with one thread.  What a warp can share:
//
//  * A pass over [beg, end) at shift s is three steps: count the digits, permute, recurse into the buckets.  Counting and
//    the bucket bookkeeping are data-parallel.  A pass whose elements all carry one digit moves nothing (ksort.h:128 finds every
//    element in its own bucket) and is skipped after the count -- with keys strand<<63 | rid<<32 | pos that is most of the
//    upper bytes.
//  * The permutation (ksort.h:126-138) is a walk: bucket d's unfilled region is a queue of ORIGINAL elements (slots at or behind
//    head[d] have never been written), the element in hand goes to the head of its digit's queue and picks up the element it
//    finds there.  So the walk needs the digits only -- one byte per element, read through L1 in queue order -- and can record
//    where every element lands (dst[]) instead of moving 16-byte elements one dependent load after the other.  One lane walks;
//    the warp then scatters a[] through dst[] with coalesced accesses.
//  * Buckets of 2..64 elements are insertion-sorted by the reference (ksort.h:143): a stable sort, one lane per bucket here.
//
// Scratch per array of n elements: tmp[n] (elements), dst[n] (int32), dig[n] (bytes), stack[n/65 + 2] frames; 512 int32 of
// shared memory per warp (bucket heads and tails).
#ifndef MMG_RSWARP_H
#define MMG_RSWARP_H
#include "mmg_warp.h"

template <class W, class T, class KeyFn>
MMG_HDN inline void mmg_rs_sort_warp(const W &wp, T *a, int64_t n, T *tmp, int32_t *dst, uint8_t *dig, RsFrame *stack,
                                     int32_t *head /* [256] */, int32_t *tail /* [256] */, KeyFn key)
{
	if (n <= 64) { // radix_sort(): rs_insertsort on the whole array -- a stable sort by key; ranks by counting
		wp.each([&](int l) {
			for (int64_t e = l; e < n; e += 32) {
				const uint64_t x = key(a[e]);
				int r = 0;
				for (int64_t q = 0; q < n; ++q) { const uint64_t y = key(a[q]); r += (y < x || (y == x && q < e)) ? 1 : 0; }
				tmp[r] = a[e];
			}
		});
		wp.each([&](int l) { for (int64_t e = l; e < n; e += 32) a[e] = tmp[e]; });
		return;
	}
	int top = 0;
	wp.one([&]() { stack[0].beg = 0, stack[0].end = n, stack[0].shift = 56; });
	top = 1;
	while (top > 0) {
		const RsFrame fr = stack[--top];
		const int s = fr.shift;
		const int32_t beg = (int32_t)fr.beg, end = (int32_t)fr.end;
		// 1. digits and their counts (ksort.h:120-121)
		wp.each([&](int l) { for (int d = l; d < 256; d += 32) tail[d] = 0; });
		wp.each([&](int l) {
			for (int32_t i = beg + l; i < end; i += 32) {
				const int d = (int)(key(a[i]) >> s & 0xff);
				dig[i] = (uint8_t)d;
				W::aadd(&tail[d], 1);
			}
		});
		// 2. bucket boundaries (ksort.h:122-123); a pass with one digit moves nothing
		int n_used = 0;
		wp.one([&]() {
			int32_t at = beg; int used = 0;
			for (int d = 0; d < 256; ++d) { const int32_t c = tail[d]; used += c != 0; head[d] = at; at += c; tail[d] = at; }
			stack[top].pad = used; // hand the count to the other lanes through memory
		});
		n_used = stack[top].pad;
		if (n_used == 1) {
			if (s) wp.one([&]() { stack[top].beg = beg, stack[top].end = end, stack[top].shift = s > 8 ? s - 8 : 0; });
			if (s) ++top;
			continue;
		}
		// 3. the permutation walk on digits (ksort.h:126-138): dst[i] = where the element now at i ends up
		wp.one([&]() {
			for (int d = 0; d < 256; ++d) {
				const int32_t td = tail[d];
				int32_t hd = head[d];
				while (hd != td) {
					int to = dig[hd];
					if (to != d) {
						int32_t carry = hd; // index of the element in hand
						do {
							const int32_t slot = head[to]; head[to] = slot + 1;
							dst[carry] = slot;
							carry = slot;   // the element found there is an original: slots >= head[to] are unwritten
							to = dig[carry];
						} while (to != d);
						dst[carry] = hd;
					} else dst[hd] = hd;
					++hd;
				}
				head[d] = hd;
			}
		});
		wp.each([&](int l) { for (int32_t i = beg + l; i < end; i += 32) tmp[dst[i]] = a[i]; });
		wp.each([&](int l) { for (int32_t i = beg + l; i < end; i += 32) a[i] = tmp[i]; });
		// 4. buckets (ksort.h:140-145): > 64 elements -> next byte; 2..64 -> insertion sort, a lane per bucket
		if (s) {
			const int ns = s > 8 ? s - 8 : 0;
			for (int d0 = 0; d0 < 256; d0 += 32) {
				const unsigned big = wp.ballot([&](int l) { const int d = d0 + l; return tail[d] - (d ? tail[d - 1] : beg) > 64; });
				wp.each([&](int l) {
					const int d = d0 + l;
					const int32_t lo = d ? tail[d - 1] : beg, hi = tail[d];
					if (hi - lo > 64) {
						RsFrame &f = stack[top + mmg_popc(big & ((1u << l) - 1u))];
						f.beg = lo, f.end = hi, f.shift = ns;
					} else if (hi - lo > 1) mmg_rs_insertion(a, (int64_t)lo, (int64_t)hi, key);
				});
				top += mmg_popc(big);
			}
		}
	}
}
#endif
