// tests/host_arith/host_arith.cpp -- CPU check of the host/device arithmetic helpers in
// abyss_b200/csrc/abb_device.cuh against the C oracle and plain operators (test infrastructure).
#include "../../abyss_b200/csrc/abb_device.cuh"
#include "../../abyss_b200/csrc/abb_walk.cuh"
#include "../../abyss_b200/host/reads.h"
extern "C" {
#include "../../oracle/abyss_oracle.h"
}
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>
using namespace abb;

static int fails = 0;
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

int main(int argc, char** argv)
{
	std::mt19937_64 rng(42);
	// exact modulo by invariant divisors, incl. the SURVEY filter sizes and adversarial numerators
	const uint64_t divs[] = { 2, 3, 8, 1000, 4096, 59652352ULL, 954437184ULL, 7635497472ULL, 61083979328ULL, (1ULL << 36) - 1,
		                      0xFFFFFFFFFFFFFFFFULL, (1ULL << 63) + 1 };
	for (uint64_t d : divs) {
		const FastMod f = make_fastmod(d);
		const uint64_t edge[] = { 0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL - d,
			                      (0xFFFFFFFFFFFFFFFFULL / d) * d, (0xFFFFFFFFFFFFFFFFULL / d) * d - 1 };
		for (uint64_t n : edge)
			CHECK(fastmod_u64(n, f) == n % d);
		for (int i = 0; i < 200000; ++i) {
			const uint64_t n = rng();
			CHECK(fastmod_u64(n, f) == n % d);
		}
	}
	// split rotation and rolls against the oracle
	for (int i = 0; i < 20000; ++i) {
		const uint64_t x = rng();
		const unsigned n = (unsigned)(rng() % 5000);
		CHECK(srol_n(x, n) == abo_srol_n(x, n));
		CHECK(sror_n(srol_n(x, n), n) == x);
		CHECK(srol1(x) == abo_srol(x) && sror1(x) == abo_sror(x));
	}
	const char* B = "ACGT";
	for (unsigned k : { 2u, 5u, 31u, 32u, 33u, 64u, 96u, 97u, 192u }) {
		std::string s(k + 300, 'A');
		for (auto& c : s)
			c = B[rng() & 3];
		const RollTab rt = make_rolltab(k);
		uint64_t fh, rh;
		abo_base_hash(s.data(), k, &fh, &rh);
		HashPair h = { fh, rh };
		Vtx<6> v = vtx_from_codes<6>((const uint8_t*)s.data(), k, true, rt);
		CHECK(v.h.fh == fh && v.h.rh == rh);
		for (unsigned i = 0; i + k < s.size(); ++i) { // roll right along the string, and check rolling back left
			abo_roll_right(&fh, &rh, k, (unsigned char)s[i], (unsigned char)s[i + k]);
			const HashPair n = roll_right(h, rt, base_code(s[i]), base_code(s[i + k]));
			CHECK(n.fh == fh && n.rh == rh);
			const HashPair back = roll_left(n, rt, base_code(s[i + k]), base_code(s[i]));
			CHECK(back.fh == h.fh && back.rh == h.rh);
			h = n;
			const unsigned out = vtx_step(v, k, rt, FWD, base_code(s[i + k]));
			CHECK(out == base_code(s[i]) && v.h.fh == fh && v.h.rh == rh);
			CHECK(kmer_first(v.km, k) == base_code(s[i + 1]) && kmer_last(v.km) == base_code(s[i + k]));
		}
		// reverse complement: hashes swap, double revcomp is the identity
		const Vtx<6> rc = vtx_revcomp(v, k);
		CHECK(rc.h.fh == v.h.rh && rc.h.rh == v.h.fh && rc.canon() == v.canon());
		const Vtx<6> rr = vtx_revcomp(rc, k);
		CHECK(kmer_equal(rr.km, v.km));
		{ // the packed reverse complement against the base-by-base definition, and isCanonical (string order)
			const std::string tail = s.substr(s.size() - k);
			std::string rcs(tail.rbegin(), tail.rend());
			for (auto& c : rcs)
				c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
			bool same = true;
			for (unsigned i = 0; i < k; ++i)
				same &= kmer_base(rc.km, k, i) == base_code(rcs[i]) && kmer_base(v.km, k, i) == base_code(tail[i]);
			CHECK(same);
			CHECK(kmer_is_canonical(v.km, k) == (tail <= rcs));
			// spaced seed: hash of the '1' positions only, from the rolled state (maskHash, nthash.hpp:417-436)
			std::string mask(k, '1');
			std::vector<uint8_t> mpos;
			for (unsigned i = 1; i + 1 < k; ++i)
				if ((i * 7 + k) % 3 == 0) {
					mask[i] = '0';
					mpos.push_back((uint8_t)i);
				}
			RollTab mrt = rt;
			mrt.nmask = (unsigned)mpos.size();
			mrt.mpos = mpos.data();
			Vtx<6> mv = v;
			vtx_rehash(mv, k, mrt);
			uint64_t hs[1];
			CHECK(abo_hash_seq(tail.data(), k, k, 1, mask.c_str(), hs, NULL) == 1 && hs[0] == mv.bloom());
			if (k > 2) {
				const std::string nxt = tail.substr(1) + "G", prv = "C" + tail.substr(0, k - 1);
				CHECK(abo_hash_seq(nxt.data(), k, k, 1, mask.c_str(), hs, NULL) == 1 && hs[0] == neighbor_bloom(mv, k, mrt, FWD, 2));
				CHECK(abo_hash_seq(prv.data(), k, k, 1, mask.c_str(), hs, NULL) == 1 && hs[0] == neighbor_bloom(mv, k, mrt, REV, 1));
			}
		}
	}
	{ // pathToSeq with the mask 10001: ACGTA, CGTAC -> ACNNAC (Unittest/BloomDBG/BloomDBGTest.cpp:24-37)
		const uint8_t mp[3] = { 1, 2, 3 };
		RollTab rt = make_rolltab(5);
		rt.nmask = 3;
		rt.mpos = mp;
		std::string out = "ACGTAC";
		for (unsigned c = 0; c < 6; ++c)
			if (!column_written(rt, 5, 2, c))
				out[c] = 'N';
		CHECK(out == "ACNNAC");
		CHECK(column_written(rt, 5, 4, 4) && column_written(rt, 5, 4, 3) && !column_written(rt, 5, 3, 3)); // k-1 vertices or more: every column written
	}
	// base codes
	CHECK(base_code('A') == 0 && base_code('c') == 1 && base_code('G') == 2 && base_code('t') == 3 && base_code('N') == 4 &&
	      base_code('U') == 4 && base_code(0) == 4);
	// SIToBytes (Common/StringUtil.h:181-219)
	uint64_t v = 0;
	CHECK(host::si_to_bytes("8G", &v) && v == (8ULL << 30));
	CHECK(host::si_to_bytes("64M", &v) && v == (64ULL << 20));
	CHECK(host::si_to_bytes("1.5k", &v) && v == 1536);
	CHECK(host::si_to_bytes("932096", &v) && v == 932096);
	CHECK(!host::si_to_bytes("8GB", &v) && !host::si_to_bytes("x", &v));
	// the FASTA/FASTQ reader on a crafted file (argv[1])
	if (argc > 1) {
		host::ReadOpts o;
		o.qualityThreshold = argc > 2 ? atoi(argv[2]) : 0;
		host::SeqReader in(argv[1], o);
		std::string id, seq;
		while (in.next(id, seq))
			printf("%s\t%s\n", id.c_str(), seq.c_str());
	}
	printf(fails ? "HOST_ARITH_FAILED %d\n" : "HOST_ARITH_OK\n", fails);
	return fails ? 1 : 0;
}
