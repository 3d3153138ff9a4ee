// abb_api.cu -- C ABI (include/abyss_b200.h): filter lifecycle, pass-1 insert, literal-hash
// interface, raw array transfer, statistics.  Host orchestration only; the arithmetic lives in
// abb_device.cuh and the kernels in abb_insert.cuh.
#include "abb_common.h"
#include "abb_insert.cuh"
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace abb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
}

constexpr uint64_t kDefaultWindow = 1ULL << 19;
constexpr uint64_t kChunkSlots = 1ULL << 25; // h0 staging: 32 Mi slots = 256 MiB + 32 MiB flags

static uint64_t next_pow2(uint64_t x)
{
	uint64_t p = 1;
	while (p < x)
		p <<= 1;
	return p;
}

int select_device(int device)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) {
		set_error("no usable CUDA device (%s); libabyssb200 has no CPU fallback",
		          e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
		return ABB_ENODEV;
	}
	if (device < 0 || device >= n) {
		set_error("device %d out of range (0..%d)", device, n - 1);
		return ABB_EINVAL;
	}
	ABB_CUDA(cudaSetDevice(device));
	// tuning knob: ABB_L2_FETCH=32 asks L2 to fetch 32 B sectors from HBM (cudaLimitMaxL2FetchGranularity).
	// Measured on B200: no gain for the random 1-byte Bloom accesses (0.85 vs 0.90 G k-mers/s), so off by default.
	static int fetch = -1;
	if (fetch < 0) {
		const char* e = getenv("ABB_L2_FETCH");
		fetch = e ? atoi(e) : 0;
	}
	if (fetch > 0)
		cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)fetch);
	return ABB_OK;
}

static FilterView view_of(const abb_filter* f)
{
	FilterView v;
	v.data = f->d_data;
	v.level_stride = f->bytes_per_level;
	v.levels = f->levels;
	return v;
}

/** make sure the ordered-insert workspace exists for the current window size: two tag tables so that
 *  the reservation pass of window w+1 overlaps the commit/resolve of window w */
static int ensure_workspace(abb_filter* f)
{
	const uint64_t want = next_pow2(2 * f->window * f->H);
	if (f->d_tags && f->tag_slots == want)
		return ABB_OK;
	if (f->d_tags)
		cudaFree(f->d_tags);
	if (f->d_deferred)
		cudaFree(f->d_deferred);
	f->d_tags = nullptr;
	f->d_deferred = nullptr;
	ABB_CUDA(cudaMalloc((void**)&f->d_tags, 2 * want * sizeof(unsigned long long)));
	ABB_CUDA(cudaMemsetAsync(f->d_tags, 0, 2 * want * sizeof(unsigned long long), f->stream));
	// two carry lists (slots that lost a reservation travel to the next window), worst case everything defers
	ABB_CUDA(cudaMalloc((void**)&f->d_deferred, 2 * (f->window + kCarryLanes) * sizeof(uint64_t)));
	f->tag_slots = want;
	f->epoch = 0;
	f->epoch2[0] = f->epoch2[1] = 0;
	if (!f->stream2) {
		ABB_CUDA(cudaStreamCreateWithFlags(&f->stream2, cudaStreamNonBlocking));
		for (int i = 0; i < 2; ++i) {
			ABB_CUDA(cudaEventCreateWithFlags(&f->ev_res[i], cudaEventDisableTiming));
			ABB_CUDA(cudaEventCreateWithFlags(&f->ev_done[i], cudaEventDisableTiming));
		}
		ABB_CUDA(cudaEventCreateWithFlags(&f->ev_in, cudaEventDisableTiming));
	}
	return ABB_OK;
}

#define ABB_DISPATCH_H(H, ...)            \
	do {                                  \
		if ((H) <= 4) {                   \
			constexpr int MAXH = 4;       \
			__VA_ARGS__;                  \
		} else if ((H) <= 8) {            \
			constexpr int MAXH = 8;       \
			__VA_ARGS__;                  \
		} else {                          \
			constexpr int MAXH = 32;      \
			__VA_ARGS__;                  \
		}                                 \
	} while (0)

/** ordered insert of slots [0, n_slots) of `hashes` (h0 per slot, or literal H per slot) */
template <bool LITERAL>
static int ordered_insert(abb_filter* f, const uint64_t* d_hashes, const uint8_t* d_valid, uint64_t n_slots)
{
	if (n_slots == 0)
		return ABB_OK;
	if (f->kind == ABB_BIT) {
		ABB_DISPATCH_H(f->H, (k_bits_insert<LITERAL, MAXH><<<blocks_for(n_slots, 256), 256, 0, f->stream>>>(
		                         d_hashes, d_valid, n_slots, f->cfg, f->d_data)));
		f->st.launches += 1;
		ABB_CUDA(cudaGetLastError());
		return ABB_OK;
	}
	ABB_CHECK(ensure_workspace(f));
	const FilterView fv = view_of(f);
	cudaStream_t s_main = f->stream, s_res = f->stream2;
	// the reservation stream may start once the hashes are there
	ABB_CUDA(cudaEventRecord(f->ev_in, s_main));
	ABB_CUDA(cudaStreamWaitEvent(s_res, f->ev_in, 0));
	const uint64_t n_windows = (n_slots + f->window - 1) / f->window;
	const unsigned age_off = (unsigned)(kAge * f->window);
	uint64_t* carry[2] = { reinterpret_cast<uint64_t*>(f->d_deferred),
		                   reinterpret_cast<uint64_t*>(f->d_deferred) + (f->window + kCarryLanes) };
	unsigned* n_carry[2] = { f->d_ndef, f->d_ndef + 1 };
	ABB_CUDA(cudaMemsetAsync(f->d_ndef, 0, 2 * sizeof(unsigned), s_main));
	bool done_recorded[2] = { false, false };
	unsigned epoch_of[2] = { 0, 0 };
	auto issue_reserve = [&](uint64_t w) -> int {
		const int p = (int)(w & 1);
		const uint64_t w0 = w * f->window;
		const unsigned n = (unsigned)std::min<uint64_t>(f->window, n_slots - w0);
		const TagTable tab = { f->d_tags + (uint64_t)p * f->tag_slots, f->tag_slots - 1 };
		if (done_recorded[p]) // table p is free again once window w-2 is through
			ABB_CUDA(cudaStreamWaitEvent(s_res, f->ev_done[p], 0));
		if (f->epoch2[p] >= kMaxEpoch) {
			ABB_CUDA(cudaMemsetAsync(tab.e, 0, f->tag_slots * sizeof(unsigned long long), s_res));
			f->epoch2[p] = 0;
		}
		epoch_of[p] = ++f->epoch2[p];
		ABB_DISPATCH_H(f->H, (k_reserve<LITERAL, MAXH><<<blocks_for(n, 256), 256, 0, s_res>>>(d_hashes, d_valid, w0, n, f->cfg, tab, epoch_of[p], age_off)));
		ABB_CUDA(cudaEventRecord(f->ev_res[p], s_res));
		return ABB_OK;
	};
	ABB_CHECK(issue_reserve(0));
	int in = 0; // carry list read by this window
	for (uint64_t w = 0; w < n_windows; ++w) {
		const int p = (int)(w & 1);
		const uint64_t w0 = w * f->window;
		const unsigned n = (unsigned)std::min<uint64_t>(f->window, n_slots - w0);
		const TagTable tab = { f->d_tags + (uint64_t)p * f->tag_slots, f->tag_slots - 1 };
		const unsigned epoch = epoch_of[p];
		if (w + 1 < n_windows)
			ABB_CHECK(issue_reserve(w + 1)); // overlaps this window's commit
		ABB_CUDA(cudaStreamWaitEvent(s_main, f->ev_res[p], 0));
		const unsigned grid = blocks_for((uint64_t)kCarryLanes + n, 256);
		// pending slots are drained when the list grows, every 8 windows (bounded age) and at the end
		const bool force = (w % 8 == 7) || (w + 1 == n_windows);
		const unsigned min_count = force ? 1u : kCarryLanes / 2;
		ABB_DISPATCH_H(f->H, {
			k_reserve_carry<LITERAL, MAXH><<<kCarryLanes / 256, 256, 0, s_main>>>(d_hashes, carry[in], n_carry[in], w0, f->cfg, tab, epoch, age_off);
			if (f->profile && f->prof_used + 2 <= f->prof_ev.size())
				cudaEventRecord(f->prof_ev[f->prof_used++], s_main);
			if (f->kind == ABB_COUNTING)
				k_commit<0, LITERAL, MAXH><<<grid, 256, 0, s_main>>>(d_hashes, d_valid, w0, n, f->cfg, tab, epoch, fv, age_off, carry[in], n_carry[in],
				                                                    carry[1 - in], n_carry[1 - in], f->d_stats);
			else
				k_commit<1, LITERAL, MAXH><<<grid, 256, 0, s_main>>>(d_hashes, d_valid, w0, n, f->cfg, tab, epoch, fv, age_off, carry[in], n_carry[in],
				                                                    carry[1 - in], n_carry[1 - in], f->d_stats);
			if (f->profile && (f->prof_used & 1))
				cudaEventRecord(f->prof_ev[f->prof_used++], s_main);
			if (f->kind == ABB_COUNTING)
				k_drain<0, LITERAL, MAXH><<<1, 1024, 0, s_main>>>(d_hashes, w0, f->cfg, tab, epoch, fv, age_off, carry[1 - in], n_carry[1 - in],
				                                                 n_carry[in], min_count, f->d_stats);
			else
				k_drain<1, LITERAL, MAXH><<<1, 1024, 0, s_main>>>(d_hashes, w0, f->cfg, tab, epoch, fv, age_off, carry[1 - in], n_carry[1 - in],
				                                                 n_carry[in], min_count, f->d_stats);
		});
		ABB_CUDA(cudaEventRecord(f->ev_done[p], s_main));
		done_recorded[p] = true;
		in = 1 - in;
		f->st.launches += 4;
		f->st.windows += 1;
	}
	ABB_CUDA(cudaGetLastError());
	if (f->profile && f->prof_used) { // fold the per-launch k_commit times into the statistics
		ABB_CUDA(cudaStreamSynchronize(s_main));
		for (size_t i = 0; i + 1 < f->prof_used; i += 2) {
			float ms = 0;
			cudaEventElapsedTime(&ms, f->prof_ev[i], f->prof_ev[i + 1]);
			f->st.ms_commit += ms;
			f->st.commit_launches += 1;
		}
		f->prof_used = 0;
	}
	return ABB_OK;
}

/** single thread: cut reads into chunks of about `cap` slots (at least one read per chunk) */
__global__ void k_chunk_bounds(const uint64_t* __restrict__ slot_offs, uint64_t n_reads, uint64_t cap,
                               uint64_t* __restrict__ bounds, unsigned max_chunks, unsigned* __restrict__ n_chunks)
{
	uint64_t r = 0;
	unsigned c = 0;
	bounds[0] = 0;
	while (r < n_reads && c + 1 < max_chunks) {
		const uint64_t limit = slot_offs[r] + cap;
		// largest r1 in (r, n_reads] with slot_offs[r1] <= limit, but at least r + 1
		uint64_t lo = r + 1, hi = n_reads;
		while (lo < hi) {
			uint64_t mid = lo + (hi - lo + 1) / 2;
			if (slot_offs[mid] <= limit)
				lo = mid;
			else
				hi = mid - 1;
		}
		r = lo;
		bounds[++c] = r;
	}
	if (r < n_reads)
		bounds[++c] = n_reads;
	*n_chunks = c;
}

__global__ void __launch_bounds__(256)
k_gather_u64(const uint64_t* __restrict__ src, const uint64_t* __restrict__ idx, uint64_t n, uint64_t* __restrict__ dst)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		dst[i] = src[idx[i]];
}

/** count valid flags (k-mers actually inserted) */
__global__ void __launch_bounds__(256)
k_count_valid(const uint8_t* __restrict__ valid, uint64_t n, unsigned long long* __restrict__ out)
{
	unsigned long long c = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		c += valid[i];
	for (int d = 16; d; d >>= 1)
		c += __shfl_down_sync(0xffffffffu, c, d);
	if ((threadIdx.x & 31) == 0 && c)
		atomicAdd(out, c);
}

/** K1 launcher for reads [r0, r1): h0/valid index = slot_offs[r] + j - slot_base */
int launch_hash(abb_filter* f, unsigned k, const uint8_t* d_care, const uint8_t* d_bases,
                       const uint64_t* d_offs, const uint64_t* d_slot_offs, uint64_t r0, uint64_t r1,
                       uint64_t slot_base, uint64_t* d_h0, uint8_t* d_valid, cudaStream_t stream, uint64_t* launches)
{
	(void)f;
	const uint64_t n = r1 - r0;
	if (n == 0)
		return ABB_OK;
	int sms = 148, dev = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const unsigned grid = (unsigned)std::min<uint64_t>((n + kHashWarps - 1) / kHashWarps, (uint64_t)sms * 32);
	if (d_care)
		k_hash_reads_masked<<<grid, kHashWarps * 32, 0, stream>>>(d_bases, d_offs + r0, d_slot_offs + r0, slot_base, n, k, d_care,
		                                                          d_h0, d_valid);
	else
		k_hash_reads<<<grid, kHashWarps * 32, 0, stream>>>(d_bases, d_offs + r0, d_slot_offs + r0, slot_base, n, k, d_h0, d_valid);
	if (launches)
		*launches += 1;
	ABB_CUDA(cudaGetLastError());
	return ABB_OK;
}

/** K1 over explicit (possibly overlapping) segments, see k_hash_segments */
int launch_hash_segments(unsigned k, const uint8_t* d_care, const uint8_t* d_bases, const uint64_t* d_seg_beg,
                         const unsigned* d_seg_len, const uint64_t* d_seg_slot, uint64_t n_segs, uint64_t* d_h0, uint8_t* d_valid,
                         cudaStream_t stream)
{
	if (n_segs == 0)
		return ABB_OK;
	int sms = 148, dev = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const unsigned grid = (unsigned)std::min<uint64_t>((n_segs + kHashWarps - 1) / kHashWarps, (uint64_t)sms * 32);
	if (d_care)
		k_hash_segments_masked<<<grid, kHashWarps * 32, 0, stream>>>(d_bases, d_seg_beg, d_seg_len, d_seg_slot, n_segs, k, d_care,
		                                                             d_h0, d_valid);
	else
		k_hash_segments<<<grid, kHashWarps * 32, 0, stream>>>(d_bases, d_seg_beg, d_seg_len, d_seg_slot, n_segs, k, d_h0, d_valid);
	ABB_CUDA(cudaGetLastError());
	return ABB_OK;
}

/** slot_offs[0..n_reads] = exclusive prefix sum of per-read window counts; returns the total */
int compute_slot_offsets(unsigned k, const uint64_t* d_offs, uint64_t n_reads, DevBuf<uint64_t>& slot_offs,
                                DevBuf<uint8_t>& tmp, cudaStream_t stream, uint64_t* total, uint64_t* launches)
{
	ABB_CHECK(slot_offs.reserve(n_reads + 1));
	ABB_CUDA(cudaMemsetAsync(slot_offs.p + n_reads, 0, sizeof(uint64_t), stream));
	if (n_reads) {
		k_window_counts<<<blocks_for(n_reads, 256), 256, 0, stream>>>(d_offs, n_reads, k, slot_offs.p);
		ABB_CUDA(cudaGetLastError());
	}
	size_t bytes = 0;
	ABB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, slot_offs.p, slot_offs.p, n_reads + 1, stream));
	ABB_CHECK(tmp.reserve(bytes));
	ABB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, slot_offs.p, slot_offs.p, n_reads + 1, stream));
	ABB_CUDA(cudaMemcpyAsync(total, slot_offs.p + n_reads, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
	ABB_CUDA(cudaStreamSynchronize(stream));
	if (launches)
		*launches += 3;
	return ABB_OK;
}

static int insert_reads_dev(abb_filter* f, const uint8_t* d_bases, const uint64_t* d_offs, uint64_t n_reads,
                            uint64_t* n_kmers_out)
{
	uint64_t total = 0;
	ABB_CHECK(compute_slot_offsets(f->k, d_offs, n_reads, f->slot_offs, f->scan_tmp, f->stream, &total, &f->st.launches));
	if (n_kmers_out)
		*n_kmers_out = 0;
	if (total == 0)
		return ABB_OK;

	// chunk the reads so that the h0 staging buffer stays bounded
	std::vector<uint64_t> bounds;
	if (total <= kChunkSlots) {
		bounds = { 0, n_reads };
	} else {
		const unsigned max_chunks = (unsigned)(total / kChunkSlots + 3) * 2;
		DevBuf<uint64_t> d_bounds;
		ABB_CHECK(d_bounds.reserve(max_chunks + 2));
		k_chunk_bounds<<<1, 1, 0, f->stream>>>(f->slot_offs.p, n_reads, kChunkSlots, d_bounds.p, max_chunks,
		                                      (unsigned*)(d_bounds.p + max_chunks + 1));
		f->st.launches += 1;
		std::vector<uint64_t> h(max_chunks + 2);
		ABB_CUDA(cudaMemcpyAsync(h.data(), d_bounds.p, (max_chunks + 2) * sizeof(uint64_t), cudaMemcpyDeviceToHost, f->stream));
		ABB_CUDA(cudaStreamSynchronize(f->stream));
		d_bounds.release();
		unsigned nc = (unsigned)(h[max_chunks + 1] & 0xffffffffu);
		bounds.assign(h.begin(), h.begin() + nc + 1);
	}
	// per-chunk slot ranges need slot_offs at the chunk boundaries
	std::vector<uint64_t> slot_at(bounds.size());
	for (size_t i = 0; i < bounds.size(); ++i)
		ABB_CUDA(cudaMemcpyAsync(&slot_at[i], f->slot_offs.p + bounds[i], sizeof(uint64_t), cudaMemcpyDeviceToHost, f->stream));
	ABB_CUDA(cudaStreamSynchronize(f->stream));

	ABB_CUDA(cudaMemsetAsync(f->d_stats + 3, 0, sizeof(unsigned long long), f->stream));
	for (size_t c = 0; c + 1 < bounds.size(); ++c) {
		const uint64_t r0 = bounds[c], r1 = bounds[c + 1];
		const uint64_t slots = slot_at[c + 1] - slot_at[c];
		if (slots == 0)
			continue;
		ABB_CHECK(f->h0.reserve(slots));
		ABB_CHECK(f->valid.reserve(slots));
		ABB_CUDA(cudaEventRecord(f->ev0, f->stream));
		ABB_CHECK(launch_hash(f, f->k, f->d_care, d_bases, d_offs, f->slot_offs.p, r0, r1, slot_at[c], f->h0.p, f->valid.p,
		                      f->stream, &f->st.launches));
		k_count_valid<<<std::min<unsigned>(blocks_for(slots, 256), 148 * 8), 256, 0, f->stream>>>(f->valid.p, slots, f->d_stats + 3);
		f->st.launches += 1;
		ABB_CUDA(cudaEventRecord(f->ev1, f->stream));
		ABB_CHECK(ordered_insert<false>(f, f->h0.p, f->valid.p, slots));
		cudaEvent_t ev2 = f->ev0; // reuse: record end of insert after reading the hash time
		ABB_CUDA(cudaEventSynchronize(f->ev1));
		float ms = 0;
		ABB_CUDA(cudaEventElapsedTime(&ms, f->ev0, f->ev1));
		f->st.ms_hash += ms;
		ABB_CUDA(cudaEventRecord(ev2, f->stream));
		ABB_CUDA(cudaEventSynchronize(ev2));
		ABB_CUDA(cudaEventElapsedTime(&ms, f->ev1, ev2));
		f->st.ms_insert += ms;
		f->st.slots += slots;
	}
	unsigned long long nk = 0;
	ABB_CUDA(cudaMemcpyAsync(&nk, f->d_stats + 3, sizeof nk, cudaMemcpyDeviceToHost, f->stream));
	ABB_CUDA(cudaStreamSynchronize(f->stream));
	f->st.kmers += nk;
	if (n_kmers_out)
		*n_kmers_out = nk;
	return ABB_OK;
}

} // namespace abb

using namespace abb;

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int abb_version(void) { return ABB_VERSION; }
const char* abb_last_error(void) { return g_err; }

int abb_device_count(void)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) {
		set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
		return ABB_ENODEV;
	}
	return n;
}

int abb_filter_create(abb_filter** out, int kind, uint64_t size, unsigned num_hashes, unsigned k, unsigned arg,
                      const char* mask, int device)
{
	ABB_REQUIRE(out != nullptr, "abb_filter_create: out is NULL");
	*out = nullptr;
	ABB_REQUIRE(kind == ABB_COUNTING || kind == ABB_BIT || kind == ABB_CASCADING, "unknown filter kind %d", kind);
	ABB_REQUIRE(num_hashes >= 1 && num_hashes <= kMaxHashes, "number of hash functions must be in 1..%u (MAX_HASHES)", kMaxHashes);
	ABB_REQUIRE(k >= 1 && k <= kMaxK, "k-mer size must be in 1..%u (MAX_KMER)", kMaxK);
	if (kind == ABB_COUNTING) {
		// CountingBloomFilter ctor pads the byte size to a multiple of 8 (CountingBloomFilter.hpp:40-49)
		if (size % 8)
			size += 8 - size % 8;
	} else {
		// BloomFilter::initSize exits on this (BloomFilter.hpp:374-379)
		ABB_REQUIRE(size % 8 == 0, "ERROR: Filter Size \"%llu\" is not a multiple of 8.", (unsigned long long)size);
	}
	ABB_REQUIRE(size >= 8, "filter size must be at least 8");
	ABB_REQUIRE(size < (1ULL << kPosBits), "filter size %llu exceeds the supported maximum 2^%u", (unsigned long long)size, kPosBits);
	unsigned levels = 1;
	if (kind == ABB_CASCADING) {
		ABB_REQUIRE(arg >= 1 && arg <= 255, "cascading filter needs 1..255 levels");
		levels = arg;
	}
	std::string m = mask ? mask : "";
	if (!m.empty()) {
		// MaskedKmer::setMask (BloomDBG/MaskedKmer.h:38-55): k long, only 0/1
		ABB_REQUIRE(m.size() == k, "spaced seed must be exactly k=%u characters long", k);
		for (char c : m)
			ABB_REQUIRE(c == '0' || c == '1', "spaced seed must contain only '0' and '1'");
		if (m.find('0') == std::string::npos)
			m.clear(); // all ones == no mask
	}
	ABB_CHECK(select_device(device));

	abb_filter* f = new (std::nothrow) abb_filter();
	if (!f) {
		set_error("out of host memory");
		return ABB_ENOMEM;
	}
	f->device = device;
	f->kind = kind;
	f->size = size;
	f->bytes_per_level = kind == ABB_COUNTING ? size : size / 8;
	f->H = num_hashes;
	f->k = k;
	f->threshold = kind == ABB_COUNTING ? arg : 0;
	f->levels = levels;
	f->mask = m;
	f->window = kDefaultWindow;
	f->cfg.H = num_hashes;
	f->cfg.k = k;
	f->cfg.mod = make_fastmod(size);
	for (unsigned i = 0; i < kMaxHashes; ++i)
		f->cfg.mult[i] = (uint64_t)i ^ ((uint64_t)k * kMultiSeed); // nthash.hpp:339
	auto fail = [&](int rc) {
		abb_filter_destroy(f);
		return rc;
	};
#define ABB_TRY(call)                                                                 \
	do {                                                                              \
		cudaError_t e__ = (call);                                                     \
		if (e__ != cudaSuccess) {                                                     \
			set_error("%s failed: %s", #call, cudaGetErrorString(e__));               \
			return fail(e__ == cudaErrorMemoryAllocation ? ABB_ENOMEM : ABB_ECUDA);   \
		}                                                                             \
	} while (0)
	ABB_TRY(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
	ABB_TRY(cudaEventCreate(&f->ev0));
	ABB_TRY(cudaEventCreate(&f->ev1));
	ABB_TRY(cudaMalloc((void**)&f->d_data, f->bytes_per_level * levels));
	ABB_TRY(cudaMemsetAsync(f->d_data, 0, f->bytes_per_level * levels, f->stream));
	ABB_TRY(cudaMalloc((void**)&f->d_ndef, 2 * sizeof(unsigned)));
	ABB_TRY(cudaMemsetAsync(f->d_ndef, 0, 2 * sizeof(unsigned), f->stream));
	ABB_TRY(cudaMalloc((void**)&f->d_stats, 8 * sizeof(unsigned long long)));
	ABB_TRY(cudaMemsetAsync(f->d_stats, 0, 8 * sizeof(unsigned long long), f->stream));
	if (!f->mask.empty()) {
		std::vector<uint8_t> care(k);
		for (unsigned i = 0; i < k; ++i)
			care[i] = f->mask[i] == '1';
		ABB_TRY(cudaMalloc((void**)&f->d_care, k));
		ABB_TRY(cudaMemcpyAsync(f->d_care, care.data(), k, cudaMemcpyHostToDevice, f->stream));
	}
	ABB_TRY(cudaStreamSynchronize(f->stream));
#undef ABB_TRY
	*out = f;
	return ABB_OK;
}

int abb_filter_destroy(abb_filter* f)
{
	if (!f)
		return ABB_OK;
	cudaSetDevice(f->device);
	if (f->stream)
		cudaStreamSynchronize(f->stream);
	cudaFree(f->d_data);
	cudaFree(f->d_care);
	cudaFree(f->d_tags);
	cudaFree(f->d_deferred);
	cudaFree(f->d_ndef);
	cudaFree(f->d_stats);
	f->bases.release();
	f->offs.release();
	f->slot_offs.release();
	f->h0.release();
	f->lit.release();
	f->valid.release();
	f->scan_tmp.release();
	f->out8.release();
	for (auto e : f->prof_ev)
		cudaEventDestroy(e);
	if (f->stream2) {
		cudaStreamSynchronize(f->stream2);
		for (int i = 0; i < 2; ++i) {
			cudaEventDestroy(f->ev_res[i]);
			cudaEventDestroy(f->ev_done[i]);
		}
		cudaEventDestroy(f->ev_in);
		cudaStreamDestroy(f->stream2);
	}
	if (f->ev0)
		cudaEventDestroy(f->ev0);
	if (f->ev1)
		cudaEventDestroy(f->ev1);
	if (f->stream)
		cudaStreamDestroy(f->stream);
	delete f;
	return ABB_OK;
}

unsigned abb_filter_kmer_size(const abb_filter* f) { return f ? f->k : 0; }
unsigned abb_filter_hash_num(const abb_filter* f) { return f ? f->H : 0; }
uint64_t abb_filter_size(const abb_filter* f) { return f ? f->size : 0; }
uint64_t abb_filter_size_in_bytes(const abb_filter* f) { return f ? f->bytes_per_level : 0; }
unsigned abb_filter_threshold(const abb_filter* f) { return f ? f->threshold : 0; }
unsigned abb_filter_levels(const abb_filter* f) { return f ? f->levels : 0; }

int abb_filter_set_threshold(abb_filter* f, unsigned threshold)
{
	ABB_REQUIRE(f, "NULL filter");
	if (f->kind != ABB_COUNTING) {
		set_error("threshold only applies to counting filters");
		return ABB_ESTATE;
	}
	f->threshold = threshold;
	return ABB_OK;
}

int abb_filter_set_profiling(abb_filter* f, int on)
{
	ABB_REQUIRE(f, "NULL filter");
	ABB_CUDA(cudaSetDevice(f->device));
	f->profile = on != 0;
	if (f->profile && f->prof_ev.empty()) {
		f->prof_ev.resize(256);
		for (auto& e : f->prof_ev)
			ABB_CUDA(cudaEventCreate(&e));
	}
	return ABB_OK;
}

int abb_filter_set_window(abb_filter* f, uint64_t window_slots)
{
	ABB_REQUIRE(f, "NULL filter");
	if (window_slots == 0)
		window_slots = kDefaultWindow;
	ABB_REQUIRE(window_slots >= 32 && window_slots <= (1ULL << 20) - 64, "window must be in [32, 2^20 - 64]");
	ABB_CUDA(cudaSetDevice(f->device));
	ABB_CUDA(cudaStreamSynchronize(f->stream));
	f->window = window_slots;
	return ABB_OK;
}

int abb_insert_reads_dev(abb_filter* f, const char* d_bases, const uint64_t* d_offsets, uint64_t n_reads,
                         uint64_t n_bases, uint64_t* n_kmers_out)
{
	(void)n_bases;
	ABB_REQUIRE(f, "NULL filter");
	ABB_REQUIRE(n_reads == 0 || (d_bases && d_offsets), "NULL read buffers");
	ABB_CUDA(cudaSetDevice(f->device));
	return insert_reads_dev(f, (const uint8_t*)d_bases, d_offsets, n_reads, n_kmers_out);
}

int abb_insert_reads(abb_filter* f, const char* bases, const uint64_t* offsets, uint64_t n_reads, uint64_t* n_kmers_out)
{
	ABB_REQUIRE(f, "NULL filter");
	if (n_kmers_out)
		*n_kmers_out = 0;
	if (n_reads == 0)
		return ABB_OK;
	ABB_REQUIRE(bases && offsets, "NULL read buffers");
	ABB_REQUIRE(offsets[0] == 0, "offsets[0] must be 0");
	const uint64_t n_bases = offsets[n_reads];
	ABB_CUDA(cudaSetDevice(f->device));
	ABB_CHECK(f->bases.reserve(n_bases + 16));
	ABB_CHECK(f->offs.reserve(n_reads + 1));
	ABB_CUDA(cudaMemcpyAsync(f->bases.p, bases, n_bases, cudaMemcpyHostToDevice, f->stream));
	ABB_CUDA(cudaMemcpyAsync(f->offs.p, offsets, (n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, f->stream));
	return insert_reads_dev(f, f->bases.p, f->offs.p, n_reads, n_kmers_out);
}

int abb_insert_hashes(abb_filter* f, const uint64_t* hashes, uint64_t n)
{
	ABB_REQUIRE(f, "NULL filter");
	if (n == 0)
		return ABB_OK;
	ABB_REQUIRE(hashes, "NULL hashes");
	ABB_CUDA(cudaSetDevice(f->device));
	ABB_CHECK(f->lit.reserve(n * f->H));
	ABB_CUDA(cudaMemcpyAsync(f->lit.p, hashes, n * f->H * sizeof(uint64_t), cudaMemcpyHostToDevice, f->stream));
	ABB_CHECK(ordered_insert<true>(f, f->lit.p, nullptr, n));
	ABB_CUDA(cudaStreamSynchronize(f->stream));
	f->st.kmers += n;
	return ABB_OK;
}

int abb_insert_h0_dev(abb_filter* f, const uint64_t* d_h0, uint64_t n)
{
	ABB_REQUIRE(f, "NULL filter");
	if (n == 0)
		return ABB_OK;
	ABB_REQUIRE(d_h0, "NULL hashes");
	ABB_CUDA(cudaSetDevice(f->device));
	ABB_CHECK(ordered_insert<false>(f, d_h0, nullptr, n));
	ABB_CUDA(cudaStreamSynchronize(f->stream));
	f->st.kmers += n;
	f->st.slots += n;
	return ABB_OK;
}

int abb_hash_reads_dev(abb_filter* f, const char* d_bases, const uint64_t* d_offsets, uint64_t n_reads, uint64_t* d_h0,
                       uint8_t* d_valid, uint64_t capacity, uint64_t* n_slots_out)
{
	ABB_REQUIRE(f, "NULL filter");
	if (n_slots_out)
		*n_slots_out = 0;
	if (n_reads == 0)
		return ABB_OK;
	ABB_REQUIRE(d_bases && d_offsets, "NULL read buffers");
	ABB_CUDA(cudaSetDevice(f->device));
	uint64_t total = 0;
	ABB_CHECK(compute_slot_offsets(f->k, d_offsets, n_reads, f->slot_offs, f->scan_tmp, f->stream, &total, &f->st.launches));
	if (n_slots_out)
		*n_slots_out = total;
	if (total == 0 || !d_h0 || !d_valid)
		return ABB_OK;
	ABB_REQUIRE(capacity >= total, "output buffers hold %llu slots, %llu needed", (unsigned long long)capacity, (unsigned long long)total);
	ABB_CHECK(launch_hash(f, f->k, f->d_care, (const uint8_t*)d_bases, d_offsets, f->slot_offs.p, 0, n_reads, 0, d_h0, d_valid, f->stream,
	                      &f->st.launches));
	ABB_CUDA(cudaStreamSynchronize(f->stream));
	return ABB_OK;
}

/** owner of a canonical hash among `world` contiguous hash ranges (abyss_b200/multigpu.py owner_of) */
struct OwnedBy {
	const uint64_t* h0;
	const uint8_t* valid;
	unsigned world, g;
	__host__ __device__ bool operator()(uint64_t i) const
	{
		return valid[i] && (unsigned)((((h0[i] >> 48) & 0xFFFF) * world) >> 16) == g;
	}
};
struct GatherH0 {
	const uint64_t* h0;
	__host__ __device__ uint64_t operator()(uint64_t i) const { return h0[i]; }
};

int abb_route_h0_dev(abb_filter* f, const uint64_t* d_h0, const uint8_t* d_valid, uint64_t n, unsigned world, uint64_t* d_send,
                     uint64_t* counts_out)
{
	ABB_REQUIRE(f && counts_out, "NULL argument");
	ABB_REQUIRE(world >= 1 && world <= 65536, "world size out of range");
	for (unsigned g = 0; g < world; ++g)
		counts_out[g] = 0;
	if (n == 0)
		return ABB_OK;
	ABB_REQUIRE(d_h0 && d_valid && d_send, "NULL buffer");
	ABB_REQUIRE(n < (1ULL << 40), "too many slots");
	ABB_CUDA(cudaSetDevice(f->device));
	// one order-preserving selection per destination: d_send = [slots owned by 0 | owned by 1 | ...]
	DevBuf<uint64_t> idx; // selected slot indices of one destination, then gathered
	unsigned long long* d_num = f->d_stats + 6;
	uint64_t at = 0;
	// select indices in pieces of < 2^31 items (cub's num_items is an int here)
	const uint64_t piece = 1ULL << 30;
	for (unsigned g = 0; g < world; ++g) {
		uint64_t got = 0;
		for (uint64_t lo = 0; lo < n; lo += piece) {
			const int m = (int)std::min<uint64_t>(piece, n - lo);
			OwnedBy pred{ d_h0, d_valid, world, g };
			thrust::counting_iterator<uint64_t> first(lo);
			size_t bytes = 0;
			ABB_CHECK(idx.reserve((size_t)m));
			ABB_CUDA(cub::DeviceSelect::If(nullptr, bytes, first, idx.p, d_num, m, pred, f->stream));
			ABB_CHECK(f->scan_tmp.reserve(bytes));
			ABB_CUDA(cub::DeviceSelect::If(f->scan_tmp.p, bytes, first, idx.p, d_num, m, pred, f->stream));
			unsigned long long k = 0;
			ABB_CUDA(cudaMemcpyAsync(&k, d_num, sizeof k, cudaMemcpyDeviceToHost, f->stream));
			ABB_CUDA(cudaStreamSynchronize(f->stream));
			if (k) {
				k_gather_u64<<<blocks_for(k, 256), 256, 0, f->stream>>>(d_h0, idx.p, k, d_send + at + got);
				ABB_CUDA(cudaGetLastError());
			}
			got += k;
			f->st.launches += 2;
		}
		counts_out[g] = got;
		at += got;
	}
	ABB_CUDA(cudaStreamSynchronize(f->stream));
	idx.release();
	return ABB_OK;
}

void* abb_filter_device_ptr(abb_filter* f, int level)
{
	if (!f)
		return nullptr;
	if (level < 0)
		level = (int)f->levels - 1;
	if ((unsigned)level >= f->levels)
		return nullptr;
	cudaSetDevice(f->device);
	cudaStreamSynchronize(f->stream);
	return f->d_data + (uint64_t)level * f->bytes_per_level;
}

static int query_hashes(abb_filter* f, const uint64_t* hashes, uint64_t n, uint8_t* out, bool want_min)
{
	ABB_REQUIRE(f, "NULL filter");
	if (n == 0)
		return ABB_OK;
	ABB_REQUIRE(hashes && out, "NULL buffer");
	ABB_CUDA(cudaSetDevice(f->device));
	ABB_CHECK(f->lit.reserve(n * f->H));
	ABB_CHECK(f->out8.reserve(n));
	ABB_CUDA(cudaMemcpyAsync(f->lit.p, hashes, n * f->H * sizeof(uint64_t), cudaMemcpyHostToDevice, f->stream));
	const FilterView fv = view_of(f);
	if (f->kind == ABB_COUNTING)
		k_query<0><<<blocks_for(n, 256), 256, 0, f->stream>>>(f->lit.p, n, f->cfg, fv, f->threshold, want_min ? nullptr : f->out8.p,
		                                                       want_min ? f->out8.p : nullptr);
	else
		k_query<1><<<blocks_for(n, 256), 256, 0, f->stream>>>(f->lit.p, n, f->cfg, fv, 0, want_min ? nullptr : f->out8.p,
		                                                       want_min ? f->out8.p : nullptr);
	f->st.launches += 1;
	ABB_CUDA(cudaGetLastError());
	ABB_CUDA(cudaMemcpyAsync(out, f->out8.p, n, cudaMemcpyDeviceToHost, f->stream));
	ABB_CUDA(cudaStreamSynchronize(f->stream));
	return ABB_OK;
}

int abb_contains_hashes(abb_filter* f, const uint64_t* hashes, uint64_t n, uint8_t* out) { return query_hashes(f, hashes, n, out, false); }
int abb_mincount_hashes(abb_filter* f, const uint64_t* hashes, uint64_t n, uint8_t* out) { return query_hashes(f, hashes, n, out, true); }

int abb_hash_reads(unsigned k, const char* mask, const char* bases, const uint64_t* offsets, uint64_t n_reads,
                   uint64_t* out_h0, uint8_t* out_valid, uint64_t* n_slots_out, int device)
{
	ABB_REQUIRE(k >= 1 && k <= kMaxK, "k-mer size must be in 1..%u", kMaxK);
	if (n_slots_out)
		*n_slots_out = 0;
	if (n_reads == 0)
		return ABB_OK;
	ABB_REQUIRE(bases && offsets, "NULL read buffers");
	ABB_CHECK(select_device(device));
	std::string m = mask ? mask : "";
	if (!m.empty()) {
		ABB_REQUIRE(m.size() == k, "spaced seed must be exactly k=%u characters long", k);
		if (m.find('0') == std::string::npos)
			m.clear();
	}
	const uint64_t n_bases = offsets[n_reads];
	DevBuf<uint8_t> d_bases, d_valid, tmp, d_care;
	DevBuf<uint64_t> d_offs, d_slot, d_h0;
	int rc = ABB_OK;
	auto cleanup = [&]() {
		d_bases.release(); d_valid.release(); tmp.release(); d_care.release();
		d_offs.release(); d_slot.release(); d_h0.release();
	};
#define ABB_TRYRC(expr) do { rc = (expr); if (rc != ABB_OK) { cleanup(); return rc; } } while (0)
	auto cu = [&](cudaError_t e, const char* what) {
		if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return (int)ABB_ECUDA; }
		return (int)ABB_OK;
	};
	ABB_TRYRC(d_bases.reserve(n_bases + 16));
	ABB_TRYRC(d_offs.reserve(n_reads + 1));
	ABB_TRYRC(cu(cudaMemcpy(d_bases.p, bases, n_bases, cudaMemcpyHostToDevice), "H2D bases"));
	ABB_TRYRC(cu(cudaMemcpy(d_offs.p, offsets, (n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice), "H2D offsets"));
	if (!m.empty()) {
		std::vector<uint8_t> care(k);
		for (unsigned i = 0; i < k; ++i)
			care[i] = m[i] == '1';
		ABB_TRYRC(d_care.reserve(k));
		ABB_TRYRC(cu(cudaMemcpy(d_care.p, care.data(), k, cudaMemcpyHostToDevice), "H2D mask"));
	}
	uint64_t total = 0;
	ABB_TRYRC(compute_slot_offsets(k, d_offs.p, n_reads, d_slot, tmp, 0, &total, nullptr));
	if (n_slots_out)
		*n_slots_out = total;
	if (total && out_h0 && out_valid) {
		ABB_TRYRC(d_h0.reserve(total));
		ABB_TRYRC(d_valid.reserve(total));
		ABB_TRYRC(launch_hash(nullptr, k, m.empty() ? nullptr : d_care.p, d_bases.p, d_offs.p, d_slot.p, 0, n_reads, 0, d_h0.p,
		                      d_valid.p, 0, nullptr));
		ABB_TRYRC(cu(cudaMemcpy(out_h0, d_h0.p, total * sizeof(uint64_t), cudaMemcpyDeviceToHost), "D2H h0"));
		ABB_TRYRC(cu(cudaMemcpy(out_valid, d_valid.p, total, cudaMemcpyDeviceToHost), "D2H valid"));
	}
#undef ABB_TRYRC
	cleanup();
	return ABB_OK;
}

static int level_ptr(abb_filter* f, int level, uint64_t nbytes, uint8_t** p)
{
	ABB_REQUIRE(f, "NULL filter");
	if (level < 0)
		level = (int)f->levels - 1;
	ABB_REQUIRE((unsigned)level < f->levels, "level %d out of range", level);
	ABB_REQUIRE(nbytes == f->bytes_per_level, "buffer is %llu bytes, the filter level is %llu", (unsigned long long)nbytes,
	            (unsigned long long)f->bytes_per_level);
	*p = f->d_data + (uint64_t)level * f->bytes_per_level;
	return ABB_OK;
}

int abb_filter_download(abb_filter* f, int level, uint8_t* host, uint64_t nbytes)
{
	uint8_t* p = nullptr;
	ABB_CHECK(level_ptr(f, level, nbytes, &p));
	ABB_REQUIRE(host, "NULL buffer");
	ABB_CUDA(cudaSetDevice(f->device));
	ABB_CUDA(cudaMemcpyAsync(host, p, nbytes, cudaMemcpyDeviceToHost, f->stream));
	ABB_CUDA(cudaStreamSynchronize(f->stream));
	return ABB_OK;
}

int abb_filter_upload(abb_filter* f, int level, const uint8_t* host, uint64_t nbytes)
{
	uint8_t* p = nullptr;
	ABB_CHECK(level_ptr(f, level, nbytes, &p));
	ABB_REQUIRE(host, "NULL buffer");
	ABB_CUDA(cudaSetDevice(f->device));
	ABB_CUDA(cudaMemcpyAsync(p, host, nbytes, cudaMemcpyHostToDevice, f->stream));
	ABB_CUDA(cudaStreamSynchronize(f->stream));
	return ABB_OK;
}

int abb_filter_clear(abb_filter* f)
{
	ABB_REQUIRE(f, "NULL filter");
	ABB_CUDA(cudaSetDevice(f->device));
	ABB_CUDA(cudaMemsetAsync(f->d_data, 0, f->bytes_per_level * f->levels, f->stream));
	ABB_CUDA(cudaStreamSynchronize(f->stream));
	return ABB_OK;
}

int abb_filter_popcount(abb_filter* f, uint64_t* nonzero, uint64_t* at_or_above_threshold)
{
	ABB_REQUIRE(f, "NULL filter");
	ABB_CUDA(cudaSetDevice(f->device));
	ABB_CUDA(cudaMemsetAsync(f->d_stats + 4, 0, 2 * sizeof(unsigned long long), f->stream));
	// bit / cascading: population of the LAST level (the one contains() consults)
	const uint8_t* p = f->d_data + (uint64_t)(f->levels - 1) * f->bytes_per_level;
	k_popcount<<<148 * 8, 256, 0, f->stream>>>(p, f->bytes_per_level, f->kind == ABB_COUNTING, f->threshold, f->d_stats + 4);
	f->st.launches += 1;
	ABB_CUDA(cudaGetLastError());
	unsigned long long h[2] = { 0, 0 };
	ABB_CUDA(cudaMemcpyAsync(h, f->d_stats + 4, sizeof h, cudaMemcpyDeviceToHost, f->stream));
	ABB_CUDA(cudaStreamSynchronize(f->stream));
	if (nonzero)
		*nonzero = h[0];
	if (at_or_above_threshold)
		*at_or_above_threshold = f->kind == ABB_COUNTING ? h[1] : h[0];
	return ABB_OK;
}

void* abb_filter_stream(abb_filter* f) { return f ? (void*)f->stream : nullptr; }

int abb_filter_insert_stats(abb_filter* f, abb_insert_stats* out, int reset)
{
	ABB_REQUIRE(f, "NULL filter");
	ABB_CUDA(cudaSetDevice(f->device));
	unsigned long long h[3] = { 0, 0, 0 };
	ABB_CUDA(cudaMemcpyAsync(h, f->d_stats, sizeof h, cudaMemcpyDeviceToHost, f->stream));
	ABB_CUDA(cudaStreamSynchronize(f->stream));
	f->st.deferred = h[0];
	if (out)
		*out = f->st;
	if (reset) {
		f->st = abb_insert_stats{};
		ABB_CUDA(cudaMemsetAsync(f->d_stats, 0, 3 * sizeof(unsigned long long), f->stream));
	}
	return ABB_OK;
}

} // extern "C"
