// common_host.cuh -- small host-side helpers shared by the evaluators inside libbcm3b200.so
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/bcm3b200.h"

namespace bcm3b200 {

inline std::string& last_error_ref()
{
	thread_local std::string e;
	return e;
}

inline int fail(int code, const char* fmt, ...)
{
	char buf[1536];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	last_error_ref() = buf;
	return code;
}

#define CUDA_TRY(expr)                                                                                                    \
	do {                                                                                                                  \
		cudaError_t e_ = (expr);                                                                                          \
		if (e_ != cudaSuccess) return bcm3b200::fail(BCM3B200_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
	} while (0)

template <class T>
struct DevBuf {
	T* p = nullptr;
	size_t n = 0;
	~DevBuf() { release(); }
	void release()
	{
		if (p) cudaFree(p);
		p = nullptr;
		n = 0;
	}
	cudaError_t ensure(size_t count)
	{
		if (count <= n) return cudaSuccess;
		release();
		cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
		if (e == cudaSuccess) n = count;
		return e;
	}
};

} // namespace bcm3b200
