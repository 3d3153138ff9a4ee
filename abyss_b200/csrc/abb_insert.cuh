// abb_insert.cuh -- pass-1 kernels: K1 hash_reads, K2 ordered (sequentially-consistent) insert.
//
// Replaces BloomDBG::loadSeq (BloomDBG/BloomIO.h:32-41) = RollingHashIterator
// (RollingHashIterator.h:35-97) + CountingBloomFilter<uint8_t>::incrementMin
// (vendor/btl_bloomfilter/CountingBloomFilter.hpp:138-162), BloomFilter::insert
// (BloomFilter.hpp:186-195) and HashAgnosticCascadingBloom::insert
// (Bloom/HashAgnosticCascadingBloom.h:124-133).
//
// WHY AN ORDERED INSERT.  incrementMin reads the minimum of H counters and bumps the counters
// equal to it, so the final array depends on the order in which k-mers that share a counter are
// inserted; the unitig FASTA prints sums of raw counters (bloom-dbg.h:599-603) and thresholds
// them, so "same FASTA as the reference at -j1" means "same counters as the sequential,
// file-order insert".  The cascading filter has the same property.  The scheme (deterministic
// reservations):
//   * k-mer windows ("slots") are numbered in file order; a batch is cut into ordered windows of
//     W slots; windows run one after another, slots inside a window run in parallel.
//   * reserve: every valid slot i writes (position, i) for each of its H filter positions into an
//     L2-resident open-addressing tag table keyed by the exact position, keeping the minimum i.
//   * commit: slot i owns position p iff the table says min == i.  A slot that owns all of its
//     positions has no earlier unfinished slot touching any of them, and no later slot can commit
//     before it (it does not own the shared position), so applying it now is what the sequential
//     order would do; owners touch pairwise disjoint positions, so plain byte loads/stores are
//     race free.  Slots that lost a reservation are appended to a deferred list.
//     Owners "release" their table entries (priority field := all ones) after applying.
//   * carry: the slots that lost a reservation are simply carried into the next window, where they
//     are the oldest pending events: priority = slot - window_start + kAge * W, so pending slots of
//     up to kAge earlier windows sort before every slot of the current window, in file order.  A
//     tiny kernel adds their reservations to the next window's table, the next commit kernel
//     gives them threads of their own.  Nothing serial remains on the per-window critical path.
//   * drain: one CTA replays whatever is still pending with the same reserve/commit/release rule
//     round by round (then strictly in order) -- every 8 windows, when the carry list grows large
//     (one k-mer repeated thousands of times), and at the end of a call.
// Table entries are [epoch:4 | position:36 | priority:24]; an entry from an older epoch is free, so
// a table is cleared only once every 15 uses.
#pragma once
#include "abb_device.cuh"
#include <cuda_runtime.h>

namespace abb {

constexpr unsigned kSlotBits = 24;                  // priority = slot - window start + kAge * W  <  (kAge + 1) * W
constexpr unsigned kAge = 15;                       // a pending slot is drained before it is this many windows old
constexpr unsigned kCarryLanes = 1u << 16;          // threads a commit launch reserves for carried slots
constexpr unsigned kPosBits = 36;                   // filters up to 2^36 counters / bits
constexpr unsigned kEpochBits = 64 - kSlotBits - kPosBits;
constexpr uint64_t kSlotMask = (1ULL << kSlotBits) - 1;
constexpr unsigned kMaxEpoch = (1u << kEpochBits) - 1; // epoch 0 = never written (memset 0)

struct TagTable {
	unsigned long long* e; // entries
	uint64_t mask;         // slots - 1 (power of two)
};

ABB_D uint64_t tag_pack(unsigned epoch, uint64_t pos, uint64_t slot)
{
	return ((uint64_t)epoch << (kPosBits + kSlotBits)) | (pos << kSlotBits) | slot;
}
ABB_D uint64_t tag_home(uint64_t pos, const TagTable& t)
{
	// Fibonacci hashing of the position
	return ((pos * 0x9E3779B97F4A7C15ULL) >> 20) & t.mask;
}

/** record "slot wants pos" keeping the smallest slot for this epoch */
ABB_D void tag_reserve(const TagTable& t, unsigned epoch, uint64_t pos, uint64_t slot)
{
	const uint64_t mine = tag_pack(epoch, pos, slot);
	const uint64_t key = mine >> kSlotBits;
	uint64_t s = tag_home(pos, t);
	for (;;) {
		unsigned long long cur = __ldcg(&t.e[s]);
		for (;;) {
			if ((cur >> (kPosBits + kSlotBits)) != epoch) { // stale or empty: claim it
				unsigned long long old = atomicCAS(&t.e[s], cur, (unsigned long long)mine);
				if (old == cur)
					return;
				cur = old;
				continue;
			}
			break;
		}
		if ((cur >> kSlotBits) == key) {
			if (mine < cur)
				atomicMin(&t.e[s], (unsigned long long)mine);
			return;
		}
		s = (s + 1) & t.mask;
	}
}

/** smallest slot that reserved pos in this epoch (pos must have been reserved) */
ABB_D uint64_t tag_owner(const TagTable& t, unsigned epoch, uint64_t pos)
{
	const uint64_t key = tag_pack(epoch, pos, 0) >> kSlotBits;
	uint64_t s = tag_home(pos, t);
	for (;;) {
		unsigned long long cur = __ldcg(&t.e[s]);
		if ((cur >> kSlotBits) == key)
			return cur & kSlotMask;
		s = (s + 1) & t.mask;
	}
}

/** like tag_owner, also reporting where the entry lives so that the release needs no second probe */
ABB_D uint64_t tag_owner_at(const TagTable& t, unsigned epoch, uint64_t pos, uint64_t* where)
{
	const uint64_t key = tag_pack(epoch, pos, 0) >> kSlotBits;
	uint64_t s = tag_home(pos, t);
	for (;;) {
		unsigned long long cur = __ldcg(&t.e[s]);
		if ((cur >> kSlotBits) == key) {
			*where = s;
			return cur & kSlotMask;
		}
		s = (s + 1) & t.mask;
	}
}

/** the owner is done with pos: later slots may now win it (slot field := kSlotMask) */
ABB_D void tag_release(const TagTable& t, unsigned epoch, uint64_t pos)
{
	const uint64_t rel = tag_pack(epoch, pos, kSlotMask);
	const uint64_t key = rel >> kSlotBits;
	uint64_t s = tag_home(pos, t);
	for (;;) {
		unsigned long long cur = __ldcg(&t.e[s]);
		if ((cur >> kSlotBits) == key) {
			__stcg(&t.e[s], (unsigned long long)rel);
			return;
		}
		s = (s + 1) & t.mask;
	}
}

// ------------------------------------------------------------------------------------------
// the three insert semantics, applied by exactly one thread that owns all H positions
// ------------------------------------------------------------------------------------------
struct FilterView {
	uint8_t* data;      // counters, or level-0 bits
	uint64_t level_stride; // bytes between cascading levels
	unsigned levels;
};

ABB_D bool bits_contain(const uint8_t* bits, const uint64_t* pos, unsigned H)
{
	bool all = true;
	for (unsigned i = 0; i < H; ++i)
		all &= (__ldcg(bits + (pos[i] >> 3)) >> (pos[i] & 7)) & 1;
	return all;
}
/** bit set by the owner of the *bit position*; neighbouring bits of the byte may belong to
 *  other owners, hence the atomic (BloomFilter.hpp:186-195 uses __sync_or_and_fetch too) */
ABB_D void bits_set(uint8_t* bits, const uint64_t* pos, unsigned H)
{
	for (unsigned i = 0; i < H; ++i) {
		uint64_t byte = pos[i] >> 3;
		unsigned* w = reinterpret_cast<unsigned*>(bits + (byte & ~3ULL));
		atomicOr(w, 1u << ((pos[i] & 7) + 8 * (byte & 3)));
	}
}
/** HashAgnosticCascadingBloom::insert (HashAgnosticCascadingBloom.h:124-133) */
ABB_D void apply_cascading(const FilterView& f, const uint64_t* pos, unsigned H)
{
	for (unsigned l = 0; l < f.levels; ++l) {
		uint8_t* bits = f.data + (uint64_t)l * f.level_stride;
		if (!bits_contain(bits, pos, H)) {
			bits_set(bits, pos, H);
			return;
		}
	}
}

// ------------------------------------------------------------------------------------------
// K1: hash_reads -- one warp per read, closed-form ntHash via warp prefix-XOR scans
// ------------------------------------------------------------------------------------------
constexpr int kHashWarps = 8;     // warps per CTA
constexpr int kRing = 256;        // per-warp ring of prefix values; needs k + 32 <= 256

ABB_D uint64_t shfl_up64(uint64_t v, int d)
{
	unsigned lo = __shfl_up_sync(0xffffffffu, (unsigned)v, d);
	unsigned hi = __shfl_up_sync(0xffffffffu, (unsigned)(v >> 32), d);
	return ((uint64_t)hi << 32) | lo;
}
ABB_D uint64_t shfl64(uint64_t v, int src)
{
	unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src);
	unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
	return ((uint64_t)hi << 32) | lo;
}

/**
 * For read r and window start j, slot = slot_offs[r] + j.
 *   h0[slot]    = canonical ntHash of bases[j .. j+k)         (RollingHash.h:69-102)
 *   valid[slot] = 1 iff the window holds only A/C/G/T         (RollingHashIterator.h:46-57)
 * P_i = XOR_{t<=i} R^{-t}(seed(c_t)), Q_i = XOR_{t<=i} R^{t}(seed(comp c_t)):
 *   fwd(j) = R^{j+k-1}(P_{j+k-1} ^ P_{j-1}),  rc(j) = R^{-j}(Q_{j+k-1} ^ Q_{j-1}).
 */
/** one warp hashes the L bases at `beg`; window j goes to slot slot0 + j */
ABB_D void hash_one_read(const uint8_t* __restrict__ bases, uint64_t beg, unsigned L, uint64_t slot0, unsigned k, uint64_t* P,
                         uint64_t* Q, unsigned* B, int lane, uint64_t* __restrict__ h0_out, uint8_t* __restrict__ valid_out)
{
	uint64_t carryP = 0, carryQ = 0;
	unsigned carryB = 0;
	for (unsigned base = 0; base < L; base += 32) {
		const unsigned i = base + lane;
		unsigned code = 4;
		if (i < L)
			code = base_code(bases[beg + i]);
		uint64_t p = 0, q = 0;
		if (code < 4) {
			p = sror_n(seed_of(code), i);
			q = srol_n(seed_of(3 - code), i);
		}
		// inclusive prefix XOR across the warp
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint64_t up = shfl_up64(p, d), uq = shfl_up64(q, d);
			if (lane >= d) {
				p ^= up;
				q ^= uq;
			}
		}
		p ^= carryP;
		q ^= carryQ;
		const unsigned badmask = __ballot_sync(0xffffffffu, code >= 4 && i < L);
		const unsigned b = carryB + __popc(badmask & (0xffffffffu >> (31 - lane)));
		P[i & (kRing - 1)] = p;
		Q[i & (kRing - 1)] = q;
		B[i & (kRing - 1)] = b;
		carryP = shfl64(p, 31);
		carryQ = shfl64(q, 31);
		carryB += __popc(badmask);
		__syncwarp();
		if (i < L && i + 1 >= k) {
			const unsigned j = i + 1 - k; // window [j, i]
			uint64_t pj = 0, qj = 0;
			unsigned bj = 0;
			if (j > 0) {
				pj = P[(j - 1) & (kRing - 1)];
				qj = Q[(j - 1) & (kRing - 1)];
				bj = B[(j - 1) & (kRing - 1)];
			}
			const uint64_t fh = srol_n(p ^ pj, i);
			const uint64_t rh = sror_n(q ^ qj, j);
			h0_out[slot0 + j] = rh < fh ? rh : fh;
			valid_out[slot0 + j] = (b == bj) ? 1 : 0;
		}
		__syncwarp();
	}
}

static __global__ void __launch_bounds__(kHashWarps * 32)
k_hash_reads(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offs,
             const uint64_t* __restrict__ slot_offs, uint64_t slot_base, uint64_t n_reads, unsigned k,
             uint64_t* __restrict__ h0_out, uint8_t* __restrict__ valid_out)
{
	__shared__ uint64_t sP[kHashWarps][kRing];
	__shared__ uint64_t sQ[kHashWarps][kRing];
	__shared__ unsigned sB[kHashWarps][kRing];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (uint64_t r = (uint64_t)blockIdx.x * kHashWarps + warp; r < n_reads; r += (uint64_t)gridDim.x * kHashWarps) {
		const uint64_t beg = offs[r];
		const unsigned L = (unsigned)(offs[r + 1] - beg);
		if (L < k)
			continue;
		hash_one_read(bases, beg, L, slot_offs[r] - slot_base, k, sP[warp], sQ[warp], sB[warp], lane, h0_out, valid_out);
	}
}

/** the same over explicit segments (long sequences are cut into overlapping pieces so that every
 *  warp has work): segment s = bases [seg_beg[s], +seg_len[s]), its first window is slot seg_slot[s] */
static __global__ void __launch_bounds__(kHashWarps * 32)
k_hash_segments(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ seg_beg, const unsigned* __restrict__ seg_len,
                const uint64_t* __restrict__ seg_slot, uint64_t n_segs, unsigned k, uint64_t* __restrict__ h0_out,
                uint8_t* __restrict__ valid_out)
{
	__shared__ uint64_t sP[kHashWarps][kRing];
	__shared__ uint64_t sQ[kHashWarps][kRing];
	__shared__ unsigned sB[kHashWarps][kRing];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (uint64_t r = (uint64_t)blockIdx.x * kHashWarps + warp; r < n_segs; r += (uint64_t)gridDim.x * kHashWarps) {
		const unsigned L = seg_len[r];
		if (L < k)
			continue;
		hash_one_read(bases, seg_beg[r], L, seg_slot[r], k, sP[warp], sQ[warp], sB[warp], lane, h0_out, valid_out);
	}
}

/** spaced-seed variant: canonical hash over the '1' positions only (maskHash, nthash.hpp:537-547;
 *  a window is bad only if a non-ACGT base sits on a '1' position, RollingHashIterator.h:58-73).
 *  One warp per read, lanes stride over windows; O(k) per window (config 4 path). */
static __global__ void __launch_bounds__(kHashWarps * 32)
k_hash_reads_masked(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offs,
                    const uint64_t* __restrict__ slot_offs, uint64_t slot_base, uint64_t n_reads, unsigned k,
                    const uint8_t* __restrict__ care /* k bytes: 1 where mask == '1' */,
                    uint64_t* __restrict__ h0_out, uint8_t* __restrict__ valid_out)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (uint64_t r = (uint64_t)blockIdx.x * kHashWarps + warp; r < n_reads;
	     r += (uint64_t)gridDim.x * kHashWarps) {
		const uint64_t beg = offs[r];
		const unsigned L = (unsigned)(offs[r + 1] - beg);
		if (L < k)
			continue;
		const uint64_t slot0 = slot_offs[r] - slot_base;
		for (unsigned j = lane; j + k <= L; j += 32) {
			uint64_t fh = 0, rh = 0;
			bool ok = true;
			for (unsigned t = 0; t < k; ++t) {
				if (!care[t])
					continue;
				unsigned code = base_code(bases[beg + j + t]);
				if (code >= 4) {
					ok = false;
					break;
				}
				fh ^= srol_n(seed_of(code), k - 1 - t);
				rh ^= srol_n(seed_of(3 - code), t);
			}
			h0_out[slot0 + j] = rh < fh ? rh : fh;
			valid_out[slot0 + j] = ok ? 1 : 0;
		}
	}
}

/** k_hash_segments for a spaced seed (unitigs of the extension stage: 'N' can only sit on '0' positions there) */
static __global__ void __launch_bounds__(kHashWarps * 32)
k_hash_segments_masked(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ seg_beg,
                       const unsigned* __restrict__ seg_len, const uint64_t* __restrict__ seg_slot, uint64_t n_segs, unsigned k,
                       const uint8_t* __restrict__ care, uint64_t* __restrict__ h0_out, uint8_t* __restrict__ valid_out)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (uint64_t r = (uint64_t)blockIdx.x * kHashWarps + warp; r < n_segs; r += (uint64_t)gridDim.x * kHashWarps) {
		const uint64_t beg = seg_beg[r];
		const unsigned L = seg_len[r];
		if (L < k)
			continue;
		const uint64_t slot0 = seg_slot[r];
		for (unsigned j = lane; j + k <= L; j += 32) {
			uint64_t fh = 0, rh = 0;
			bool ok = true;
			for (unsigned t = 0; t < k; ++t) {
				if (!care[t])
					continue;
				unsigned code = base_code(bases[beg + j + t]);
				if (code >= 4) {
					ok = false;
					break;
				}
				fh ^= srol_n(seed_of(code), k - 1 - t);
				rh ^= srol_n(seed_of(3 - code), t);
			}
			h0_out[slot0 + j] = rh < fh ? rh : fh;
			valid_out[slot0 + j] = ok ? 1 : 0;
		}
	}
}

/** per-read window counts -> (exclusive scan done by the caller with cub-free two-pass code) */
static __global__ void k_window_counts(const uint64_t* __restrict__ offs, uint64_t n_reads, unsigned k,
                                uint64_t* __restrict__ counts)
{
	uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r < n_reads) {
		uint64_t L = offs[r + 1] - offs[r];
		counts[r] = L >= k ? L - k + 1 : 0;
	}
}

// ------------------------------------------------------------------------------------------
// K2: ordered insert over one window of slots [w0, w0 + n)
// ------------------------------------------------------------------------------------------
constexpr unsigned kMaxRounds = 48;

/** positions of slot s: either derived from h0 (stride 1) or read from the literal
 *  H-per-k-mer array of the reference interface */
template <bool LITERAL, int MAXH>
ABB_D void slot_positions(const uint64_t* __restrict__ hashes, uint64_t s, const HashCfg& cfg, uint64_t* pos)
{
	if (LITERAL) {
#pragma unroll
		for (int i = 0; i < MAXH; ++i)
			if (i < (int)cfg.H)
				pos[i] = fastmod_u64(hashes[s * cfg.H + i], cfg.mod);
	} else {
		const uint64_t h0 = hashes[s];
#pragma unroll
		for (int i = 0; i < MAXH; ++i)
			if (i < (int)cfg.H)
				pos[i] = nth_pos(h0, cfg, i);
	}
}

/** CountingBloomFilter::incrementMin / HashAgnosticCascadingBloom::insert by the single owner */
template <int KIND, int MAXH>
ABB_D void apply_owner(const FilterView& f, const uint64_t* pos, unsigned H)
{
	if (KIND == 0) {
		unsigned v[MAXH];
		unsigned mn = 255;
#pragma unroll
		for (int i = 0; i < MAXH; ++i)
			if (i < (int)H) {
				v[i] = __ldcg(f.data + pos[i]);
				mn = min(mn, v[i]);
			}
		if (mn == 255) // "if (minVal > newVal) return": saturated (CountingBloomFilter.hpp:146-149)
			return;
#pragma unroll
		for (int i = 0; i < MAXH; ++i)
			if (i < (int)H && v[i] == mn)
				__stcg(f.data + pos[i], (uint8_t)(mn + 1));
	} else
		apply_cascading(f, pos, H);
}

/** K2a: reserve */
template <bool LITERAL, int MAXH>
__global__ void __launch_bounds__(256)
k_reserve(const uint64_t* __restrict__ hashes, const uint8_t* __restrict__ valid, uint64_t w0,
          unsigned n, HashCfg cfg, TagTable tab, unsigned epoch, unsigned age_off)
{
	const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n)
		return;
	const uint64_t s = w0 + t;
	if (valid && !valid[s])
		return;
	uint64_t pos[MAXH];
	slot_positions<LITERAL, MAXH>(hashes, s, cfg, pos);
#pragma unroll
	for (int i = 0; i < MAXH; ++i)
		if (i < (int)cfg.H)
			tag_reserve(tab, epoch, pos[i], age_off + t);
}

/** reservations of the slots carried over from earlier windows (they precede every slot of this one) */
template <bool LITERAL, int MAXH>
__global__ void __launch_bounds__(256)
k_reserve_carry(const uint64_t* __restrict__ hashes, const uint64_t* __restrict__ carry, const unsigned* __restrict__ n_carry,
                uint64_t w0, HashCfg cfg, TagTable tab, unsigned epoch, unsigned age_off)
{
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= *n_carry)
		return;
	const uint64_t s = carry[i];
	uint64_t pos[MAXH];
	slot_positions<LITERAL, MAXH>(hashes, s, cfg, pos);
	const uint64_t prio = s + age_off - w0;
#pragma unroll
	for (int j = 0; j < MAXH; ++j)
		if (j < (int)cfg.H)
			tag_reserve(tab, epoch, pos[j], prio);
}

/** K2b: owners apply and release, everybody else is carried into the next window.
 *  Threads [0, kCarryLanes) serve the slots carried in from earlier windows, the rest the n slots
 *  of this window.  The counter loads are issued before the ownership probes (wasted only for the
 *  <1 % of slots that lose a reservation) so that the HBM round trip overlaps the L2 round trip. */
template <int KIND, bool LITERAL, int MAXH>
__global__ void __launch_bounds__(256)
k_commit(const uint64_t* __restrict__ hashes, const uint8_t* __restrict__ valid, uint64_t w0,
         unsigned n, HashCfg cfg, TagTable tab, unsigned epoch, FilterView f, unsigned age_off,
         const uint64_t* __restrict__ carry_in, const unsigned* __restrict__ n_in,
         uint64_t* __restrict__ carry_out, unsigned* __restrict__ n_out, unsigned long long* __restrict__ stats)
{
	const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t s;
	if (id < kCarryLanes) {
		if (id >= *n_in)
			return;
		s = carry_in[id];
	} else {
		const unsigned t = id - kCarryLanes;
		if (t >= n)
			return;
		s = w0 + t;
		if (valid && !valid[s])
			return;
	}
	const uint64_t prio = s + age_off - w0;
	uint64_t pos[MAXH];
	slot_positions<LITERAL, MAXH>(hashes, s, cfg, pos);
	unsigned v[MAXH];
	if (KIND == 0) {
#pragma unroll
		for (int i = 0; i < MAXH; ++i)
			if (i < (int)cfg.H)
				v[i] = __ldcg(f.data + pos[i]);
	}
	bool owner = true;
	uint64_t where[MAXH];
#pragma unroll
	for (int i = 0; i < MAXH; ++i)
		if (i < (int)cfg.H)
			owner &= tag_owner_at(tab, epoch, pos[i], &where[i]) == prio;
	if (owner) {
		if (KIND == 0) {
			// CountingBloomFilter::incrementMin by the single owner (CountingBloomFilter.hpp:138-162)
			unsigned mn = 255;
#pragma unroll
			for (int i = 0; i < MAXH; ++i)
				if (i < (int)cfg.H)
					mn = min(mn, v[i]);
			if (mn != 255) {
#pragma unroll
				for (int i = 0; i < MAXH; ++i)
					if (i < (int)cfg.H && v[i] == mn)
						__stcg(f.data + pos[i], (uint8_t)(mn + 1));
			}
		} else
			apply_owner<KIND, MAXH>(f, pos, cfg.H);
#pragma unroll
		for (int i = 0; i < MAXH; ++i)
			if (i < (int)cfg.H)
				__stcg(&tab.e[where[i]], (unsigned long long)tag_pack(epoch, pos[i], kSlotMask));
	} else {
		carry_out[atomicAdd(n_out, 1u)] = s;
		if (id >= kCarryLanes)
			atomicAdd(&stats[0], 1ULL); // slots that did not commit in their own window
	}
}

/** K2c: one CTA applies every pending slot of `list` in dependency order (see the header comment).
 *  Does nothing when fewer than min_count slots are pending.  Clears *n_consumed (the list the
 *  preceding commit launch has just read) so that it can collect the next window's carries.
 *  stats[1] = max rounds seen, stats[2] += slots replayed strictly in order. */
template <int KIND, bool LITERAL, int MAXH>
__global__ void __launch_bounds__(1024)
k_drain(const uint64_t* __restrict__ hashes, uint64_t w0, HashCfg cfg, TagTable tab, unsigned epoch, FilterView f, unsigned age_off,
        uint64_t* __restrict__ list, unsigned* __restrict__ n_list, unsigned* __restrict__ n_consumed, unsigned min_count,
        unsigned long long* __restrict__ stats)
{
	__shared__ unsigned s_left;
	__shared__ unsigned long long s_key[32];
	if (threadIdx.x == 0)
		*n_consumed = 0;
	const unsigned n = *n_list;
	if (n < min_count || n == 0)
		return;
	constexpr uint64_t kDone = ~0ULL;
	unsigned left = n, round = 0;
	while (left > 0 && round < kMaxRounds) {
		for (unsigned x = threadIdx.x; x < n; x += blockDim.x) {
			const uint64_t s = list[x];
			if (s == kDone)
				continue;
			uint64_t pos[MAXH];
			slot_positions<LITERAL, MAXH>(hashes, s, cfg, pos);
#pragma unroll
			for (int i = 0; i < MAXH; ++i)
				if (i < (int)cfg.H)
					tag_reserve(tab, epoch, pos[i], s + age_off - w0);
		}
		if (threadIdx.x == 0)
			s_left = 0;
		__threadfence();
		__syncthreads();
		for (unsigned x = threadIdx.x; x < n; x += blockDim.x) {
			const uint64_t s = list[x];
			if (s == kDone)
				continue;
			uint64_t pos[MAXH];
			slot_positions<LITERAL, MAXH>(hashes, s, cfg, pos);
			bool owner = true;
#pragma unroll
			for (int i = 0; i < MAXH; ++i)
				if (i < (int)cfg.H)
					owner &= tag_owner(tab, epoch, pos[i]) == s + age_off - w0;
			if (owner) {
				apply_owner<KIND, MAXH>(f, pos, cfg.H);
#pragma unroll
				for (int i = 0; i < MAXH; ++i)
					if (i < (int)cfg.H)
						tag_release(tab, epoch, pos[i]);
				list[x] = kDone;
			} else
				atomicAdd(&s_left, 1u);
		}
		__threadfence();
		__syncthreads();
		left = s_left;
		++round;
		__syncthreads();
	}
	const unsigned serial = left;
	// strict in-order replay of whatever is left (only reachable through very long chains)
	while (left > 0) {
		unsigned long long best = ~0ULL;
		unsigned bestx = 0;
		for (unsigned x = threadIdx.x; x < n; x += blockDim.x) {
			const uint64_t s = list[x];
			if (s < best) {
				best = s;
				bestx = x;
			}
		}
		// block-wide argmin on (slot, x): slots fit in 40 bits
		unsigned long long key = best == ~0ULL ? ~0ULL : ((best << 24) | bestx);
		for (int d = 16; d; d >>= 1) {
			unsigned long long o = __shfl_down_sync(0xffffffffu, key, d);
			key = o < key ? o : key;
		}
		if ((threadIdx.x & 31) == 0)
			s_key[threadIdx.x >> 5] = key;
		__syncthreads();
		if (threadIdx.x == 0) {
			unsigned long long m = s_key[0];
			for (unsigned w = 1; w < (blockDim.x >> 5); ++w)
				m = s_key[w] < m ? s_key[w] : m;
			const uint64_t s = m >> 24;
			const unsigned x = (unsigned)(m & 0xffffff);
			uint64_t pos[MAXH];
			slot_positions<LITERAL, MAXH>(hashes, s, cfg, pos);
			apply_owner<KIND, MAXH>(f, pos, cfg.H);
			list[x] = kDone;
		}
		__threadfence();
		__syncthreads();
		--left;
	}
	if (threadIdx.x == 0) {
		*n_list = 0;
		atomicMax(&stats[1], (unsigned long long)round);
		atomicAdd(&stats[2], (unsigned long long)serial);
	}
}

/** BloomFilter::insert for every valid slot (order free: OR commutes) -- the assembled-k-mer
 *  filter and plain `abyss-bloom build` bit filters */
template <bool LITERAL, int MAXH>
__global__ void __launch_bounds__(256)
k_bits_insert(const uint64_t* __restrict__ hashes, const uint8_t* __restrict__ valid, uint64_t n,
              HashCfg cfg, uint8_t* __restrict__ bits)
{
	const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n || (valid && !valid[s]))
		return;
	uint64_t pos[MAXH];
	slot_positions<LITERAL, MAXH>(hashes, s, cfg, pos);
	bits_set(bits, pos, cfg.H);
}

/** contains()/minCount() for the literal hash interface */
template <int KIND>
__global__ void __launch_bounds__(256)
k_query(const uint64_t* __restrict__ hashes, uint64_t n, HashCfg cfg, FilterView f, unsigned threshold,
        uint8_t* __restrict__ out_contains, uint8_t* __restrict__ out_min)
{
	const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n)
		return;
	if (KIND == 0) {
		unsigned mn = 255;
		for (unsigned i = 0; i < cfg.H; ++i)
			mn = min(mn, (unsigned)__ldcg(f.data + fastmod_u64(hashes[s * cfg.H + i], cfg.mod)));
		if (out_min)
			out_min[s] = (uint8_t)mn;
		if (out_contains)
			out_contains[s] = mn >= threshold;
	} else {
		// bit filter, or last level of a cascading filter (HashAgnosticCascadingBloom.h:105-109)
		const uint8_t* bits = f.data + (uint64_t)(f.levels - 1) * f.level_stride;
		uint64_t pos[kMaxHashes];
		for (unsigned i = 0; i < cfg.H; ++i)
			pos[i] = fastmod_u64(hashes[s * cfg.H + i], cfg.mod);
		bool c = bits_contain(bits, pos, cfg.H);
		if (out_contains)
			out_contains[s] = c;
		if (out_min)
			out_min[s] = c;
	}
}

/** popCount / filtered_popcount (CountingBloomFilter.hpp:219-244) and getPop (BloomFilter.hpp:313-320) */
static __global__ void __launch_bounds__(256)
k_popcount(const uint8_t* __restrict__ data, uint64_t nbytes, int counting, unsigned threshold,
           unsigned long long* __restrict__ out /* [2] */)
{
	unsigned long long nz = 0, th = 0;
	const uint64_t nwords = nbytes / 16;
	const uint4* w = reinterpret_cast<const uint4*>(data);
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords;
	     i += (uint64_t)gridDim.x * blockDim.x) {
		uint4 v = w[i];
		unsigned x[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			if (counting) {
#pragma unroll
				for (int b = 0; b < 4; ++b) {
					unsigned c = (x[j] >> (8 * b)) & 0xff;
					nz += c != 0;
					th += c >= threshold;
				}
			} else
				nz += __popc(x[j]);
		}
	}
	// tail bytes
	if (blockIdx.x == 0 && threadIdx.x == 0)
		for (uint64_t i = nwords * 16; i < nbytes; ++i) {
			unsigned c = data[i];
			if (counting) {
				nz += c != 0;
				th += c >= threshold;
			} else
				nz += __popc(c);
		}
	for (int d = 16; d; d >>= 1) {
		nz += __shfl_down_sync(0xffffffffu, nz, d);
		th += __shfl_down_sync(0xffffffffu, th, d);
	}
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(&out[0], nz);
		atomicAdd(&out[1], th);
	}
}

} // namespace abb
