/*
 * oracle.h — prototypes of the CPU oracle (TEST INFRASTRUCTURE; see oracle.c header).
 * Same op/dtype ids as the product ABI (include/agpu.h) so tests can drive both with one
 * table; host pointers everywhere; validity is handled by separate functions exactly as the
 * reference does it (a value kernel plus a bitmap kernel per op).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include "../include/agpu.h"

#ifdef __cplusplus
extern "C" {
#endif

int oracle_num_threads(void);
void oracle_set_num_threads(int n);
/* 1 = reproduce reference bugs Q3 (signed u32 min/max) and Q7 (merge validity with a
 * missing operand bitmap); default 0 */
void oracle_set_ref_quirks(int on);

int oracle_validity_and(const uint32_t* va, const uint32_t* vb, uint32_t* vout, size_t n_bits);
int oracle_binary(int op, int dtype, const void* a, const void* b, void* out, size_t n);
int oracle_scalar(int op, int dtype, const void* a, const void* scalar, void* out, size_t n);
int oracle_unary(int op, int dtype, const void* a, void* out, size_t n);
int oracle_compare(int op, int dtype, const void* a, const void* b, uint32_t* out_bits, size_t n);
int oracle_shift(int op, int dtype, const void* a, const uint32_t* counts, void* out, size_t n);
int oracle_bitmap_binary(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n_bits);
int oracle_bitmap_not(const uint32_t* a, uint32_t* out, size_t n_bits);
int oracle_cast(int src_dtype, int dst_dtype, const void* a, void* out, size_t n);
int oracle_merge(int dtype, const void* a, const void* b, const uint32_t* mask, void* out, size_t n);
int oracle_merge_validity(const uint32_t* va, const uint32_t* vb, const uint32_t* mask,
                          const uint32_t* vmask, uint32_t* vout, size_t n);
int oracle_take(int dtype, const void* src, size_t src_len, const uint32_t* idx, void* out, size_t m);
int oracle_put(int dtype, const void* src, const uint32_t* src_idx, void* dst,
               const uint32_t* dst_idx, size_t m);
int oracle_filter(int dtype, const void* src, const uint32_t* vsrc, const uint32_t* mask,
                  const uint32_t* vmask, size_t n, void* out, uint32_t* vout, uint64_t* count);
int oracle_broadcast(int dtype, const void* scalar, void* out, size_t n);
int oracle_sum(int dtype, const void* a, size_t n, void* out);
int oracle_any(const uint32_t* bits, size_t n_bits, uint32_t* result);
int oracle_all(const uint32_t* bits, size_t n_bits, uint32_t* result);

#ifdef __cplusplus
}
#endif
#endif
