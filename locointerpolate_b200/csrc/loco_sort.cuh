// loco_sort.cuh -- the record format of leaf-sorted queries (written by scatter_kernel, read by the tile kernels)
#pragma once
#include "loco_basis.cuh"

namespace loco {

// Sorted query record: N cube coordinates followed by the original query index, padded to whole 32-byte
// sectors and written with 256-bit stores, so that the random scatter only ever writes FULL sectors (a
// partial-sector write makes the L2 fetch the rest of the sector from HBM first).
__host__ __device__ inline int sorted_row_doubles(int N) { return ((N * 8 + 4 + 31) / 32) * 4; }
__device__ __forceinline__ void st_sector(double *p, double a, double b, double c, double d)
{
	asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld_sector(const double *p, double &a, double &b, double &c, double &d)
{
	asm volatile("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
template <int NT>
__device__ __forceinline__ void store_sorted_row(double *row, const Coord<NT> &x, int32_t qi, int N)
{
	const int RS = sorted_row_doubles(N);
	double v[(NT > 0 ? ((NT * 8 + 4 + 31) / 32) * 4 : ((kMaxN * 8 + 4 + 31) / 32) * 4)];
#pragma unroll
	for (int d = 0; d < RS; d++) v[d] = d < N ? x.get(d) : 0.0;
	v[RS - 1] = __longlong_as_double((long long)qi);
#pragma unroll
	for (int d = 0; d < RS; d += 4) st_sector(row + d, v[d], v[d + 1], v[d + 2], v[d + 3]);
}
template <int NT>
__device__ __forceinline__ int32_t load_sorted_row(const double *row, Coord<NT> &x, int N)
{
	const int RS = sorted_row_doubles(N);
	double v[(NT > 0 ? ((NT * 8 + 4 + 31) / 32) * 4 : ((kMaxN * 8 + 4 + 31) / 32) * 4)];
#pragma unroll
	for (int d = 0; d < RS; d += 4) ld_sector(row + d, v[d], v[d + 1], v[d + 2], v[d + 3]);
#pragma unroll
	for (int d = 0; d < N; d++) x.set(d, v[d]);
	return (int32_t)__double_as_longlong(v[RS - 1]);
}


} // namespace loco
