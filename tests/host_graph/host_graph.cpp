// tests/host_graph/host_graph.cpp -- TEST INFRASTRUCTURE ONLY (never linked into libabyssb200).
//
// CPU emulation of `abyss-bloom-dbg -g`: the product's traversal (abyss_b200/host/graph_dump.h) with the two GPU queries
// replaced by single-thread loops over the SAME device functions (successors_chain of csrc/abb_graph.cuh, the hash helpers of
// abb_device.cuh) on a counting filter built by the C oracle.  Output is compared with the unmodified reference's -g file.
//
//   host_graph K KC H COUNTERS reads.fq > graph.dot
#include "../../abyss_b200/csrc/abb_graph.cuh"
#include "../../abyss_b200/host/graph_dump.h"
extern "C" {
#include "../../oracle/abyss_oracle.h"
}

using namespace abb;

int main(int argc, char** argv)
{
	if (argc < 6) {
		fprintf(stderr, "usage: host_graph K KC H COUNTERS reads.fq\n");
		return 2;
	}
	const unsigned k = atoi(argv[1]), kc = atoi(argv[2]), H = atoi(argv[3]);
	const uint64_t m = strtoull(argv[4], 0, 10);
	const std::vector<std::string> files{ argv[5] };
	host::ReadOpts ropt;
	std::vector<uint8_t> counters(m, 0);
	{
		host::SeqReader in(files[0], ropt);
		std::string id, seq;
		while (in.next(id, seq))
			abo_cbf_load_seq(counters.data(), m, seq.data(), seq.size(), k, H, nullptr);
	}
	HashCfg cfg;
	cfg.H = H;
	cfg.k = k;
	cfg.mod = make_fastmod(m);
	for (unsigned i = 0; i < kMaxHashes; ++i)
		cfg.mult[i] = (uint64_t)i ^ ((uint64_t)k * kMultiSeed);
	auto probe = [&](uint64_t h0) {
		for (unsigned i = 0; i < H; ++i)
			if (counters[nth_pos(h0, cfg, i)] < kc)
				return false;
		return true;
	};
	host::output_graph(
	    k, 0,
	    [&](auto fn) {
		    host::BatchStream stream(files, ropt, 700, 2);
		    while (const host::ReadBatch* b = stream.next())
			    fn(*b);
	    },
	    [&](const char* bases, const uint64_t* offsets, uint64_t n, uint8_t* flag, uint8_t* valid, uint64_t) {
		    uint64_t s = 0;
		    std::vector<uint64_t> h(H * 4096);
		    std::vector<uint32_t> pos(4096);
		    for (uint64_t r = 0; r < n; ++r) {
			    const char* seq = bases + offsets[r];
			    const uint64_t L = offsets[r + 1] - offsets[r];
			    const uint64_t w = L >= k ? L - k + 1 : 0;
			    for (uint64_t p = 0; p < w; ++p)
				    valid[s + p] = flag[s + p] = 0;
			    if (w) {
				    if (h.size() < w * H) {
					    h.resize(w * H);
					    pos.resize(w);
				    }
				    const size_t nv = abo_hash_seq(seq, L, k, H, nullptr, h.data(), pos.data());
				    for (size_t i = 0; i < nv; ++i) {
					    valid[s + pos[i]] = 1;
					    flag[s + pos[i]] = probe(h[i * H]); // h[i*H] is the canonical hash; the others follow from it
				    }
			    }
			    s += w;
		    }
	    },
	    [&](const char* kmers, uint64_t n, unsigned max_chain, abb_succ_info* info, unsigned* len, uint64_t* self) {
		    for (uint64_t i = 0; i < n; ++i) {
			    for (unsigned s = 0; s < max_chain; ++s)
				    info[i * max_chain + s] = abb_succ_info{};
			    len[i] = successors_chain((const uint8_t*)kmers + i * k, k, max_chain, probe, info + i * max_chain, self + i);
		    }
	    },
	    std::cout);
	return 0;
}
