// The Hessian block sink shared by the contact-row kernel (barrier_kernels.cu) and the element kernels of the other terms of
// the Newton system (barrier_kernels.cu: flow / mass; elastic_kernels.cu: membrane / hinge).
#pragma once

namespace idp {

// Hessian sink: block (i,j), i <= j, of a row goes to the bucket of its lower vertex vlo = min(v[i], v[j]) (stored transposed
// when v[i] > v[j]). A row reserves its slots in the (up to four) buckets with one atomic each; inside the reservation the
// blocks are ordered by the stencil index of the higher vertex. Bucket entry: key (vhi << 32 | row << 4 | 4 i + j) -- the
// low word is a deterministic origin tag that fixes the summation order of duplicates -- and the 3x3 values split 64 + 8
// bytes (the 64-byte part is written as two full 32-byte sectors).
struct BucketEmit {
    unsigned long long* key; double* val8; double* val1;
    int base[4]; int nv; int v[4]; unsigned rowTag; // v / base are only indexed with compile-time constants (registers, not local memory)
    __device__ __forceinline__ void reserve(int* cursor)
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            base[k] = 0;
            if (k < nv) {
                int m = 1;
#pragma unroll
                for (int j = 0; j < 4; ++j) m += (j < nv && v[j] > v[k]) ? 1 : 0;
                base[k] = atomicAdd(cursor + v[k], m); // the cursors start at the bucket offsets: nothing depends on the result before the first block is stored
            }
        }
    }
    __device__ __forceinline__ bool wants(int i, int j) const { return i <= j; }
    // i, j are compile-time constants at every call site (unrolled loops)
    __device__ __forceinline__ void operator()(int i, int j, const double* blk) const
    {
        const bool tr = v[i] > v[j];
        const int va = tr ? v[j] : v[i], vb = tr ? v[i] : v[j]; // va: lower vertex, vb: higher (equal on the diagonal)
        const int ba = tr ? base[j] : base[i];
        int rnk = 0; // blocks of the row in bucket va are ordered by the stencil index of the higher vertex
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const bool before = tr ? (m < i) : (m < j);
            const bool isA = tr ? (m == j) : (m == i);
            rnk += (before && m < nv && (v[m] > va || isA)) ? 1 : 0;
        }
        const long s = (long)ba + rnk;
        key[s] = ((unsigned long long)(unsigned)vb << 32) | (unsigned long long)(rowTag | (unsigned)(4 * i + j));
        double t[9];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) t[3 * p + q] = tr ? blk[3 * q + p] : blk[3 * p + q];
        double2* d8 = reinterpret_cast<double2*>(val8 + 8 * s);
        d8[0] = make_double2(t[0], t[1]); d8[1] = make_double2(t[2], t[3]); d8[2] = make_double2(t[4], t[5]); d8[3] = make_double2(t[6], t[7]);
        val1[s] = t[8];
    }
};

} // namespace idp
