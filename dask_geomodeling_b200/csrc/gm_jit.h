// Internal interface of the specialising evaluator back end (gm_jit.cu).
#pragma once
#include "gm_common.cuh"

namespace gm {

// 1 when every lookup table of `prog` is baked into the generated module (no
// device copy of the tables is needed by jit_launch).
int jit_tables_baked(const GmProgram* prog);

// Compile (once per distinct program) and launch the specialised kernel over
// device pointers.  Table pointers are device pointers, used only for tables
// that are not baked.
int jit_launch(const GmProgram* prog, const void* const* in, const int* in_dtype, void* const* out,
               const int* out_dtype, const void* const* tkeys, const void* const* tvals,
               const void* const* thit, int64_t n, cudaStream_t s);

int64_t jit_compile_count();

}  // namespace gm
