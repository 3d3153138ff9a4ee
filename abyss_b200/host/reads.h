// reads.h -- host-side FASTA/FASTQ ingestion for the B200 CLI.
//
// Mirrors the semantics of the reference reader for the formats the Bloom-DBG stage is fed with
// (DataLayer/FastaReader.cpp:130-421, FOLD_CASE flag; Common/Uncompress.cpp for .gz/.bz2/.xz):
//   '>' FASTA (multi-line) and '@' FASTQ records, '#' comment lines, Casava 1.8 headers with the
//   chastity filter (opt::chastityFilter, default on), trimming of masked (lower-case) ends
//   (opt::trimMasked, default on -- FastaReader.cpp:29), quality trimming (-q) and masking (-Q),
//   quality offset 33/64, case folding.  SAM/qseq/export and colour-space input are not handled
//   (the reference's FastaReader.cpp:270-360 paths): the CLI reports them as unsupported.
#pragma once
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace host {

struct ReadOpts {
	int chastityFilter = 1;
	int trimMasked = 1;
	int qualityThreshold = 0;
	int internalQThreshold = 0;
	int qualityOffset = 0; // 0 = format default (33)
};

/** a batch of reads: concatenated bases + offsets, ids kept for the FASTA comments */
struct ReadBatch {
	std::vector<char> bases;
	std::vector<uint64_t> offsets{ 0 };
	std::vector<std::string> ids;
	void clear()
	{
		bases.clear();
		offsets.assign(1, 0);
		ids.clear();
	}
	size_t size() const { return offsets.size() - 1; }
	void add(const std::string& id, const std::string& seq)
	{
		ids.push_back(id);
		bases.insert(bases.end(), seq.begin(), seq.end());
		offsets.push_back(bases.size());
	}
};

class SeqReader {
  public:
	SeqReader(const std::string& path, const ReadOpts& o) : m_path(path), m_opt(o)
	{
		const char* p = path.c_str();
		auto ends = [&](const char* suf) {
			size_t n = strlen(suf);
			return path.size() >= n && path.compare(path.size() - n, n, suf) == 0;
		};
		std::string cmd;
		if (ends(".gz") || ends(".z") || ends(".Z"))
			cmd = "gunzip -c '" + path + "'";
		else if (ends(".bz2"))
			cmd = "bunzip2 -c '" + path + "'";
		else if (ends(".xz"))
			cmd = "xzdec -c '" + path + "'";
		if (!cmd.empty()) {
			m_f = popen(cmd.c_str(), "r");
			m_pipe = true;
		} else if (path == "-")
			m_f = stdin;
		else
			m_f = fopen(p, "r");
		if (!m_f) {
			fprintf(stderr, "error: `%s': %s\n", p, strerror(errno)); // assert_good (Common/IOUtil.h:14-22)
			exit(EXIT_FAILURE);
		}
		m_buf.resize(1 << 22);
		setvbuf(m_f, m_buf.data(), _IOFBF, m_buf.size());
	}
	~SeqReader()
	{
		if (m_f && m_f != stdin)
			m_pipe ? pclose(m_f) : fclose(m_f);
	}

	/** next record; false at end of file */
	bool next(std::string& id, std::string& seq)
	{
		std::string comment, q, line;
		for (;;) {
			int c = peek();
			while (c == '#') { // discard comments
				getline(line);
				c = peek();
			}
			if (c == EOF)
				return false;
			if (c != '>' && c != '@') {
				getline(line);
				die();
				fprintf(stderr, "only FASTA ('>') and FASTQ ('@') input is supported by the B200 CLI, saw `%c' near\n%s\n", c, line.c_str());
				exit(EXIT_FAILURE);
			}
			std::string header;
			getline(header);
			if (header.size() > 3 && header[0] == '@' && isalpha(header[1]) && isalpha(header[2]) && header[3] == '\t')
				continue; // SAM header line
			const char type = header[0];
			size_t i = 1;
			while (i < header.size() && !isspace((unsigned char)header[i]))
				++i;
			id = header.substr(1, i - 1);
			while (i < header.size() && isspace((unsigned char)header[i]))
				++i;
			comment = header.substr(i);
			bool skip = false;
			if (comment.size() > 3 && comment[1] == ':' && comment[3] == ':') { // Casava: read:chastity:flags:index
				if (m_opt.chastityFilter && comment[2] == 'Y')
					skip = true;
				else if (id.size() > 2 && id[id.size() - 2] != '/') {
					id += '/';
					id += comment[0];
				}
			}
			getline(seq);
			if (type == '>') {
				for (int p = peek(); p != '>' && p != '#' && p != EOF; p = peek()) {
					getline(line);
					seq += line;
				}
				q.clear();
			} else {
				int plus = getc(m_f);
				if (plus != '+') {
					die();
					fprintf(stderr, "expected `+' and saw `%c'\n", plus);
					exit(EXIT_FAILURE);
				}
				getline(line);
				getline(q);
			}
			if (skip)
				continue;
			if (seq.empty()) {
				die();
				fprintf(stderr, "sequence with ID `%s' is empty\n", id.c_str());
				exit(EXIT_FAILURE);
			}
			if (!q.empty() && q.size() != seq.size()) {
				die();
				fprintf(stderr, "sequence and quality must be the same length near\n%s\n%s\n", seq.c_str(), q.c_str());
				exit(EXIT_FAILURE);
			}
			if (m_opt.trimMasked) { // FastaReader.cpp:236-250
				size_t front = 0, back = seq.size();
				while (front < seq.size() && islower((unsigned char)seq[front]))
					++front;
				while (back > 0 && islower((unsigned char)seq[back - 1]))
					--back;
				if (front >= back) {
					seq.clear();
					q.clear();
				} else {
					seq = seq.substr(front, back - front);
					if (!q.empty())
						q = q.substr(front, back - front);
				}
			}
			for (auto& ch : seq) // FOLD_CASE
				ch = (char)toupper((unsigned char)ch);
			const int qoff = m_opt.qualityOffset > 0 ? m_opt.qualityOffset : 33;
			if (m_opt.qualityThreshold > 0 && !q.empty()) { // FastaReader.cpp:376-394
				const int good = qoff + m_opt.qualityThreshold;
				size_t front = 0, back = q.size();
				while (front < q.size() && (unsigned char)q[front] < good)
					++front;
				while (back > 0 && (unsigned char)q[back - 1] < good)
					--back;
				if (front >= back) {
					seq.erase(1);
					q.erase(1);
				} else {
					seq = seq.substr(front, back - front);
					q = q.substr(front, back - front);
				}
			}
			if (m_opt.internalQThreshold > 0 && !q.empty()) { // FastaReader.cpp:396-407
				const int good = qoff + m_opt.internalQThreshold;
				for (size_t j = 0; j < q.size(); ++j)
					if ((unsigned char)q[j] < good)
						seq[j] = 'N';
			}
			return true;
		}
	}

  private:
	FILE* die()
	{
		fprintf(stderr, "%s:%llu: error: ", m_path.c_str(), (unsigned long long)m_line);
		return stderr;
	}
	int peek()
	{
		int c = getc(m_f);
		if (c != EOF)
			ungetc(c, m_f);
		return c;
	}
	bool getline(std::string& s)
	{
		s.clear();
		char* line = nullptr;
		size_t cap = 0;
		ssize_t n = ::getline(&line, &cap, m_f);
		if (n < 0) {
			free(line);
			return false;
		}
		++m_line;
		while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r'))
			--n;
		s.assign(line, (size_t)n);
		free(line);
		return true;
	}
	std::string m_path;
	ReadOpts m_opt;
	FILE* m_f = nullptr;
	bool m_pipe = false;
	std::vector<char> m_buf;
	uint64_t m_line = 0;
};

/** SIToBytes (Common/StringUtil.h:181-219): number with optional k/M/G suffix (powers of 1024) */
inline bool si_to_bytes(const char* s, uint64_t* out)
{
	char* end = nullptr;
	double v = strtod(s, &end);
	if (end == s)
		return false;
	if (*end) {
		if (end[1])
			return false;
		switch (tolower((unsigned char)*end)) {
		case 'k': v *= (double)(1ULL << 10); break;
		case 'm': v *= (double)(1ULL << 20); break;
		case 'g': v *= (double)(1ULL << 30); break;
		default: return false;
		}
	}
	*out = (uint64_t)std::ceil(v);
	return true;
}

} // namespace host
