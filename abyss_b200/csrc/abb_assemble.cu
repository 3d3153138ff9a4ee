// abb_assemble.cu -- pass 2 (BloomDBG::assemble / processRead) behind the C ABI.
// Work in progress: entry points exist so the ABI is complete; the kernels land next.
#include "abb_common.h"

struct abb_assembler {
	abb_filter* solid = nullptr;
	abb_assembly_params params = {};
	abb_assembly_counters counters = {};
};

extern "C" {

int abb_assembler_create(abb_assembler** out, abb_filter* solid, const abb_assembly_params* params)
{
	ABB_REQUIRE(out && solid && params, "NULL argument");
	*out = nullptr;
	abb::set_error("abb_assembler_create: pass 2 is not implemented in this build");
	return ABB_ESTATE;
}
int abb_assembler_destroy(abb_assembler* a) { delete a; return ABB_OK; }
int abb_assembler_process_reads(abb_assembler*, const char*, const uint64_t*, uint64_t, const abb_contig**, uint64_t*, const char**)
{
	abb::set_error("pass 2 is not implemented in this build");
	return ABB_ESTATE;
}
int abb_assembler_counters(const abb_assembler* a, abb_assembly_counters* out)
{
	ABB_REQUIRE(a && out, "NULL argument");
	*out = a->counters;
	return ABB_OK;
}
int abb_assembler_read_results(const abb_assembler*, const uint8_t**, uint64_t*)
{
	abb::set_error("pass 2 is not implemented in this build");
	return ABB_ESTATE;
}
abb_filter* abb_assembler_assembled_filter(abb_assembler*) { return nullptr; }

} // extern "C"
