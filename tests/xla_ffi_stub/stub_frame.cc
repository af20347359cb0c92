// C entry points that let a Python test fill the test double's call frame (see xla/ffi/api/ffi.h in this directory).
#include "xla/ffi/api/ffi.h"

using xla::ffi::stub::CallFrame;
using xla::ffi::stub::RawBuffer;

extern "C" {
CallFrame* ust_ffi_stub_frame_new(void* stream) { auto* f = new CallFrame(); f->stream = stream; return f; }
void ust_ffi_stub_frame_free(CallFrame* f) { delete f; }
static RawBuffer make(void* data, int dtype, int rank, const int64_t* dims) { return RawBuffer{data, dtype, std::vector<int64_t>(dims, dims + rank)}; }
void ust_ffi_stub_add_arg(CallFrame* f, void* data, int dtype, int rank, const int64_t* dims) { f->args.push_back(make(data, dtype, rank, dims)); }
void ust_ffi_stub_add_ret(CallFrame* f, void* data, int dtype, int rank, const int64_t* dims) { f->rets.push_back(make(data, dtype, rank, dims)); }
void ust_ffi_stub_set_attr(CallFrame* f, const char* name, double v) { f->attrs[name] = v; }
const char* ust_ffi_stub_error(CallFrame* f) { return f->error.c_str(); }
}
