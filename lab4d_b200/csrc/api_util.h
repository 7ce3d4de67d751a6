// Shared helpers of the C-ABI translation units (api.cu, api_train.cu).
#pragma once
#include <string>

#include "handle.h"

static inline int fail(b200r_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
static inline int fail_cuda(b200r_handle* h, cudaError_t e, const char* where) {
  return fail(h, B200R_E_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
