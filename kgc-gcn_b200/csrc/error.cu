// Thread-local error text + ABI version for libkgc_b200.so.
#include "common.cuh"

namespace kgc {
namespace {
thread_local std::string g_last_error;
}
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace kgc

extern "C" const char* kgc_last_error(void) { return kgc::g_last_error.c_str(); }
extern "C" int kgc_abi_version(void) { return KGC_ABI_VERSION; }
