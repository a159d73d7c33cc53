/*
 * agpu.h — C ABI of the B200-native columnar compute path ("libagpu.so").
 *
 * This is the drop-in boundary for psvri/arrow-gpu's hot path.  Every entry point below
 * replaces one seam of the reference's wgpu layer; the reference file:line it replaces is
 * cited per function (paths relative to the reference root).  The reference's operator
 * crates (arithmetic, compare, logical, cast, math, trigonometry, routines) call
 *     ArrowComputePipeline::apply_{unary,binary,ternary,scalar,broadcast}_function
 *         (crates/array/src/gpu_utils/compute_pipeline.rs:24-256)
 * with a (WGSL source, entry point) pair; here the pair becomes an (op id, dtype id) pair
 * and the wgpu::Buffer arguments become raw device pointers.  INTEGRATION.md shows the
 * `extern "C"` block and `build.rs` the Rust crates would add.
 *
 * Conventions
 *  - plain C: pointers and sizes only, no C++/torch types; every function returns
 *    0 on success, a positive cudaError_t, or a negative AGPU_E* code.  Nothing here
 *    falls back to the CPU: without a CUDA device every compute call fails.
 *  - value buffers hold `n` tightly packed little-endian elements of the dtype
 *    (crates/array/src/array/primitive_array_gpu.rs:12-19).  Bitmaps (boolean data and
 *    validity) are LSB-first bits packed in uint32 words, ceil(n/32) words long, bit i =
 *    word i/32, mask 1<<(i%32) (== byte i/8, mask 1<<(i%8);
 *    crates/array/src/array/null_bit_buffer.rs:47-49).  Validity 1 = valid.  A NULL
 *    validity pointer means "no bitmap = all valid" (null_bit_buffer.rs:99-111).
 *  - kernels read each input once and write each output once; bits >= n of every bitmap
 *    word written are zero.
 *  - all work is enqueued on the device handle's stream and returns without host
 *    synchronisation (like ArrowComputePipeline::finish, compute_pipeline.rs:259-273);
 *    only agpu_d2h / agpu_sync / agpu_event_elapsed_ms wait.
 *  - fast paths need 16-byte aligned pointers (anything from agpu_alloc is 256-byte
 *    aligned); other alignments take a slower element-wise kernel with identical results.
 */
#ifndef AGPU_H
#define AGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGPU_ABI_VERSION 2

/* dtype ids — crates/array/src/array/mod.rs:40-50 (enum ArrowType) */
typedef enum {
  AGPU_BOOL = 0,
  AGPU_I8 = 1,
  AGPU_I16 = 2,
  AGPU_I32 = 3,
  AGPU_U8 = 4,
  AGPU_U16 = 5,
  AGPU_U32 = 6,
  AGPU_F32 = 7,
  AGPU_DATE32 = 8 /* i32 storage, crates/array/src/array/date32_gpu.rs */
} agpu_dtype;

/* binary ops (array∘array and array∘scalar) */
typedef enum {
  AGPU_ADD = 0, /* arithmetic/compute_shaders/{f32,i32,u32}/array.wgsl, scalar.wgsl */
  AGPU_SUB = 1,
  AGPU_MUL = 2,
  AGPU_DIV = 3,
  AGPU_REM = 4,
  AGPU_MIN = 5, /* compare/compute_shaders/ * /min_max.wgsl */
  AGPU_MAX = 6,
  AGPU_AND = 7, /* logical/compute_shaders/{i32,u32}/logical.wgsl */
  AGPU_OR = 8,
  AGPU_XOR = 9,
  AGPU_POW = 10 /* math/compute_shaders/f32/floatbinary.wgsl, i32/binary.wgsl */
} agpu_binop;

/* unary ops */
typedef enum {
  AGPU_NEG = 0,   /* arithmetic/compute_shaders/f32/neg.wgsl */
  AGPU_ABS = 1,   /* math f32/floatunary.wgsl:42, i32/unary.wgsl */
  AGPU_NOT = 2,   /* logical {i32,u32}/not.wgsl */
  AGPU_SQRT = 3,  /* math f32/floatunary.wgsl */
  AGPU_CBRT = 4,
  AGPU_EXP = 5,
  AGPU_EXP2 = 6,
  AGPU_LOG = 7,
  AGPU_LOG2 = 8,
  AGPU_SIN = 9,   /* trigonometry * /trigonometry.wgsl ; ints: fused cast->f32 */
  AGPU_COS = 10,
  AGPU_ACOS = 11,
  AGPU_SINH = 12  /* trigonometry * /hyperbolic.wgsl */
} agpu_unop;

/* compare ops — compare/compute_shaders/ * /cmp.wgsl */
typedef enum { AGPU_GT = 0, AGPU_GTEQ = 1, AGPU_LT = 2, AGPU_LTEQ = 3, AGPU_EQ = 4 } agpu_cmpop;

/* shift ops — logical/compute_shaders/ * /shift.wgsl */
typedef enum { AGPU_SHL = 0, AGPU_SHR = 1 } agpu_shiftop;

/* error codes (negative; positive values are cudaError_t) */
#define AGPU_IPC_HANDLE_BYTES 64 /* sizeof(cudaIpcMemHandle_t) */
#define AGPU_MAX_SHARDS 16       /* GPUs of one box a column can be sharded over */
#define AGPU_OK 0
#define AGPU_EUNSUPPORTED (-1) /* op/dtype pair the path does not define */
#define AGPU_EINVAL (-2)       /* NULL where a pointer is required, bad size ... */
#define AGPU_ENODEVICE (-3)    /* no CUDA device / device handle invalid */
#define AGPU_EDOUBLEFREE (-4)  /* agpu_free of a block that is already free / was never allocated here */
#define AGPU_ETIMEOUT (-5)     /* a peer GPU did not post its value (agpu_exchange_wait) */

typedef struct agpu_device agpu_device; /* opaque: {ordinal, cudaStream_t, cudaMemPool_t} */
typedef struct agpu_event agpu_event;   /* cudaEvent_t wrapper; its first member IS the cudaEvent_t, so an agpu_event* can be
                                           used as the `cudaEvent_t* sync_event` of an Arrow C Device Data Interface
                                           ArrowDeviceArray, and such a sync_event can be passed to agpu_stream_wait_event */
typedef struct agpu_graph agpu_graph;   /* opaque: a captured pipeline (cudaGraphExec_t + its temporaries) */

/* ---- device / buffer layer: replaces GpuDevice (crates/array/src/gpu_utils/gpu_device.rs) ---- */

/* GpuDevice::new, gpu_device.rs:46-85 — opens CUDA device `ordinal`, creates one
 * non-blocking stream and a stream-ordered memory pool that retains freed blocks. */
int agpu_device_create(int ordinal, agpu_device** out);
int agpu_device_destroy(agpu_device* dev);
/* number of visible CUDA devices (0 when there is no driver/GPU) */
int agpu_device_count(int* out);
/* the cudaStream_t all work of this handle is ordered on (as void*) */
void* agpu_device_stream(agpu_device* dev);
int agpu_device_ordinal(agpu_device* dev);
/* kernels launched by this handle since creation (bench.py's gpu_launches) */
uint64_t agpu_launch_count(agpu_device* dev);
const char* agpu_error_string(int code);
int agpu_abi_version(void);

/* create_empty_buffer, gpu_device.rs:183-192 — stream-ordered, NOT zero-filled.  A pointer may
 * only be used by work ordered after the allocation on this handle's stream (or on another
 * stream that waits for it), and must be freed through the same handle. */
int agpu_alloc(agpu_device* dev, size_t bytes, void** out);
/* Drop of wgpu::Buffer — stream-ordered free, safe right after enqueueing work.  Freed blocks
 * are cached by the handle that allocated them and reused by its later agpu_alloc calls of a
 * similar size.  `dev` may be any handle of the same GPU: the block goes back to its owner, but
 * only after `dev` and every handle recorded with agpu_buffer_record_use have got past the work
 * they had enqueued at this moment (one event per handle; nothing waits, the block is just not
 * handed out before).  Freeing twice (or a pointer agpu_alloc never returned) fails with
 * AGPU_EDOUBLEFREE and changes nothing. */
int agpu_free(agpu_device* dev, void* ptr);
/* wgpu keeps a buffer alive until every submitted command that uses it has finished
 * (buffer.rs:5-7: Arc<wgpu::Buffer>).  With one stream per handle that is automatic for the
 * allocating handle; when ANOTHER handle `dev` (e.g. the compute handle reading a column that an
 * upload handle allocated) enqueues work on the block at `ptr`, it says so here, and after
 * agpu_free the block is not handed out again before that handle's stream got there. */
int agpu_buffer_record_use(agpu_device* dev, const void* ptr);
/* return every cached free block to the driver */
int agpu_trim(agpu_device* dev);
/* create_gpu_buffer_with_data, gpu_device.rs:171-181 (host may be pageable or pinned) */
int agpu_h2d(agpu_device* dev, void* dst_dev, const void* src_host, size_t bytes);
/* retrive_data, gpu_device.rs:232-265 — copies and waits for the stream */
int agpu_d2h(agpu_device* dev, void* dst_host, const void* src_dev, size_t bytes);
/* same copy without the wait (dst_host should be pinned); ordered on the handle's stream */
int agpu_d2h_async(agpu_device* dev, void* dst_host, const void* src_dev, size_t bytes);
/* clone_buffer / copy_buffer_to_buffer, gpu_device.rs:212-230, compute_pipeline.rs:275-300 */
int agpu_d2d(agpu_device* dev, void* dst_dev, const void* src_dev, size_t bytes);
int agpu_memset(agpu_device* dev, void* dst_dev, int byte_value, size_t bytes);
/* device.poll(Wait) */
int agpu_sync(agpu_device* dev);
/* pinned host staging for from_slice / raw_values */
int agpu_host_alloc(size_t bytes, void** out);
int agpu_host_free(void* ptr);

/* CmpQuery timestamp queries, crates/array/src/gpu_utils/compute_query.rs:3-90 */
int agpu_event_create(agpu_event** out);
int agpu_event_destroy(agpu_event* ev);
int agpu_event_record(agpu_device* dev, agpu_event* ev);
int agpu_event_elapsed_ms(agpu_event* start, agpu_event* stop, float* ms); /* waits for stop */
/* make all later work of `dev` wait (on the GPU, not the host) for an event recorded on another
 * handle's stream: lets an upload stream run ahead of the compute stream */
int agpu_stream_wait_event(agpu_device* dev, agpu_event* ev);

/* ---- one submit per recorded pipeline ----
 * ArrowComputePipeline records every op into one wgpu CommandEncoder and `finish()` submits the
 * whole program at once (compute_pipeline.rs:259-273).  CUDA analogue: everything enqueued on the
 * handle between agpu_graph_begin and agpu_graph_end is captured into a CUDA graph instead of
 * running; agpu_graph_launch replays all of it with ONE driver call (programmatic dependent launch
 * edges between the streaming kernels included).  Buffers allocated inside the capture keep their
 * addresses: those still alive belong to the caller (the replay overwrites them), those freed
 * inside the capture stay with the graph until agpu_graph_destroy.  Host-synchronising calls
 * (agpu_d2h, agpu_sync) are not allowed while capturing. */
int agpu_graph_begin(agpu_device* dev);
int agpu_graph_end(agpu_device* dev, agpu_graph** out);
int agpu_graph_launch(agpu_device* dev, agpu_graph* graph);
uint64_t agpu_graph_kernel_count(agpu_graph* graph); /* kernels one replay launches */
int agpu_graph_destroy(agpu_graph* graph);

/* ---- validity bitmaps: NullBitBufferGpu (crates/array/src/array/null_bit_buffer.rs) ---- */

/* merge_null_bit_buffer[_op], null_bit_buffer.rs:168-243: vout = va & vb; a NULL side
 * means all-valid (then vout = copy of the other side).  Both NULL -> AGPU_EINVAL
 * (the reference returns None without touching the GPU). */
int agpu_validity_and(agpu_device* dev, const uint32_t* va, const uint32_t* vb, uint32_t* vout,
                      size_t n_bits);

/* ---- arithmetic / min-max / logical / power, array ∘ array ----
 * arithmetic/src/lib.rs:54-94 (impl_arithmetic_array_op), compare/src/lib.rs:113-140,164-172,
 * logical/src/lib.rs:88-131, math/src/lib.rs:203-209.
 * out[i] = a[i] op b[i]; if vout != NULL also vout = va & vb in the same pass. */
int agpu_binary(agpu_device* dev, int op, int dtype, const void* a, const void* b, void* out,
                size_t n, const uint32_t* va, const uint32_t* vb, uint32_t* vout);

/* ---- array ∘ scalar: arithmetic/src/lib.rs:11-50 (impl_arithmetic_op) ----
 * `scalar_dev` points at a 1-element device array, as in the reference where the rhs of
 * *_scalar is a length-1 PrimitiveArrayGpu.  Validity is copied (lib.rs:35-38). */
int agpu_scalar(agpu_device* dev, int op, int dtype, const void* a, const void* scalar_dev,
                void* out, size_t n, const uint32_t* va, uint32_t* vout);

/* ---- unary: neg / abs / not / f32 math / trig ----
 * arithmetic_kernels.rs:296-319, math/src/lib.rs:195-237, logical/src/lib.rs:133-158,
 * trigonometry/src/lib.rs:115-137.  For SIN/COS/SINH on i8/u8/i16/u16 the output is f32
 * (cast fused, trigonometry/compute_shaders/{i8,u8,i16,u16}); otherwise out has `dtype`. */
int agpu_unary(agpu_device* dev, int op, int dtype, const void* a, void* out, size_t n,
               const uint32_t* va, uint32_t* vout);

/* ---- compare -> packed bitmap: compare/src/lib.rs:85-111,142-162 ----
 * bit i of out_bits = a[i] op b[i]; ceil(n/32) words written, padding bits zero. */
int agpu_compare(agpu_device* dev, int op, int dtype, const void* a, const void* b,
                 uint32_t* out_bits, size_t n, const uint32_t* va, const uint32_t* vb,
                 uint32_t* vout);

/* ---- shifts: logical/src/lib.rs:160-186; counts is a UInt32 column, one per row ---- */
int agpu_shift(agpu_device* dev, int op, int dtype, const void* a, const uint32_t* counts,
               void* out, size_t n, const uint32_t* va, const uint32_t* vcounts, uint32_t* vout);

/* ---- bitmap (BooleanArrayGPU) logical: logical/src/boolean.rs:45-75 ----
 * op in {AGPU_AND, AGPU_OR, AGPU_XOR}; padding bits of the last word are cleared. */
int agpu_bitmap_binary(agpu_device* dev, int op, const uint32_t* a, const uint32_t* b,
                       uint32_t* out, size_t n_bits, const uint32_t* va, const uint32_t* vb,
                       uint32_t* vout);
int agpu_bitmap_not(agpu_device* dev, const uint32_t* a, uint32_t* out, size_t n_bits,
                    const uint32_t* va, uint32_t* vout);

/* ---- casts: cast/src/lib.rs:40-87,135-161 (matrix), boolean_cast.rs, f32_cast.rs ----
 * src/dst dtype pairs outside the reference matrix return AGPU_EUNSUPPORTED.
 * For src == AGPU_BOOL `a` is a bitmap.  Same-width casts are copies, and so is the one
 * bitcast the reference has, BitCast<Float32ArrayGPU> for UInt32ArrayGPU (cast/src/lib.rs:90-108,
 * 187-192), requested here as (AGPU_U32 -> AGPU_F32): the bits are reinterpreted, not converted. */
int agpu_cast(agpu_device* dev, int src_dtype, int dst_dtype, const void* a, void* out,
              size_t n, const uint32_t* va, uint32_t* vout);

/* ---- fused expression  ((a*b)+c) > d  on f32 columns (BASELINE.json config 3) ----
 * One pass instead of mul_op_dyn -> add_op_dyn -> gt_op_dyn on a shared pipeline
 * (crates/arrow/examples/simple.rs:45-72 pattern).  Bit-identical to the unfused chain:
 * two roundings, no FMA contraction.  vout = va & vb & vc & vd (NULL = all valid). */
int agpu_fused_mul_add_gt(agpu_device* dev, const float* a, const float* b, const float* c,
                          const float* d, uint32_t* out_bits, size_t n, const uint32_t* va,
                          const uint32_t* vb, const uint32_t* vc, const uint32_t* vd,
                          uint32_t* vout);

/* ---- general fused LINEAR chains on f32 (north_star (2): "cast and f32 math/trig are fused where
 * chained") ----
 * acc = f32(in[i]); then for each step, in order:
 *   AGPU_STEP_UNARY          acc = unop(acc)                 (NEG ABS SQRT CBRT EXP EXP2 LOG LOG2 SIN COS ACOS SINH)
 *   AGPU_STEP_BINARY_COLUMN  acc = acc binop operand[i]      (ADD SUB MUL DIV REM MIN MAX POW; operand = f32 column)
 *   AGPU_STEP_BINARY_SCALAR  acc = acc binop scalar
 *   AGPU_STEP_COMPARE_COLUMN / _SCALAR   bit = acc cmpop rhs (only as the LAST step; out is then a bitmap)
 * Every step rounds exactly like the stand-alone kernel of that op, so a fused chain is
 * bit-identical to the same ops recorded one by one on an ArrowComputePipeline
 * (crates/arrow/examples/simple.rs:45-72) — it just reads each input once and writes one output.
 * in_dtype: F32, I8, U8, I16, U16 (the int->f32 cast is fused like trigonometry/compute_shaders/
 * {i8,u8,i16,u16}).  At most AGPU_CHAIN_MAX_STEPS steps and 3 operand columns.  vout = AND of
 * the validity bitmaps of the input and of the operand columns (NULL = all valid). */
#define AGPU_CHAIN_MAX_STEPS 8
typedef enum {
  AGPU_STEP_UNARY = 0,
  AGPU_STEP_BINARY_COLUMN = 1,
  AGPU_STEP_BINARY_SCALAR = 2,
  AGPU_STEP_COMPARE_COLUMN = 3,
  AGPU_STEP_COMPARE_SCALAR = 4,
  /* rhs = a ONE-element f32 array on the device (how the reference passes scalars:
   * arithmetic/src/lib.rs:11-50); `operand` points at it, no host round trip */
  AGPU_STEP_BINARY_DEVSCALAR = 5,
  AGPU_STEP_COMPARE_DEVSCALAR = 6,
  /* agpu_fused_chain_int only: acc = acc << / >> counts[i]; op = AGPU_SHL / AGPU_SHR, operand = a
   * u32 column of per-row counts (logical/src/lib.rs:160-186), validity = its bitmap */
  AGPU_STEP_SHIFT_COLUMN = 7,
  /* agpu_fused_chain_pair only: out_value[i] = acc / acc = f32(in[i]) again */
  AGPU_STEP_STORE = 8,
  AGPU_STEP_RESET = 9
} agpu_step_kind;
typedef struct {
  int32_t kind;             /* agpu_step_kind */
  int32_t op;               /* agpu_unop / agpu_binop / agpu_cmpop id */
  const void* operand;      /* device column for *_COLUMN steps / one-element array for *_DEVSCALAR steps
                             * (f32 for agpu_fused_chain, the column type for agpu_fused_chain_int) */
  const uint32_t* validity; /* its validity bitmap or NULL */
  float scalar;             /* immediate for *_SCALAR steps */
} agpu_chain_step;
int agpu_fused_chain(agpu_device* dev, int in_dtype, const void* in, const uint32_t* vin,
                     const agpu_chain_step* steps, int n_steps, void* out, size_t n, uint32_t* vout);

/* TWO results of one source column in ONE pass: a value chain and a predicate chain, e.g. the
 * reference's first benchmark program  s = a + b;  g = a > b  (BASELINE.json configs[0]) reads a
 * and b once instead of twice and is one launch instead of two.  steps =
 *   value-chain steps, AGPU_STEP_STORE, AGPU_STEP_RESET, predicate-chain steps (last one a compare)
 * in_dtype F32 only; the steps are the arithmetic ones (NEG ABS SQRT, ADD SUB MUL DIV REM MIN MAX,
 * compares), <= AGPU_CHAIN_MAX_STEPS including STORE and RESET, <= 3 DISTINCT operand columns (a
 * column used by both chains counts once and is loaded once).  Every step rounds like the
 * stand-alone kernel, so both results are bit-identical to the two chains run separately.  ONE
 * validity bitmap is written: vout = AND of vin and of every operand column's bitmap — the caller
 * pairs only chains whose validity inputs are the same set. */
int agpu_fused_chain_pair(agpu_device* dev, int in_dtype, const void* in, const uint32_t* vin,
                          const agpu_chain_step* steps, int n_steps, float* out_value, uint32_t* out_bits,
                          size_t n, uint32_t* vout);

/* The same chain machinery on INTEGER columns (dtype = I8 U8 I16 U16 I32 U32 DATE32): the running
 * value, every operand column and every device scalar have type `dtype`, and each step is the
 * stand-alone integer kernel of that op — wrap in the column's own width, x/0 = x, x%0 = 0,
 * MIN/-1 = MIN, signedness of min/max/compare — so the result is bit-identical to the ops run one
 * by one (logical/src/lib.rs:120-158, arithmetic/src/lib.rs:11-94, compare/src/lib.rs:142-172).
 *   AGPU_STEP_UNARY              NOT (ABS for I32)
 *   AGPU_STEP_BINARY_COLUMN / _DEVSCALAR   ADD SUB MUL DIV REM MIN MAX AND OR XOR (POW for I32)
 *   AGPU_STEP_SHIFT_COLUMN                 SHL SHR by a u32 counts column (count & 31 on the widened lane);
 *                                          one shift step per chain, columns + counts <= 3
 *   AGPU_STEP_COMPARE_COLUMN / _DEVSCALAR  GT GTEQ LT LTEQ EQ, last step only; out is a bitmap
 * Immediates (*_SCALAR steps) are not accepted: the float field cannot hold every 32-bit integer. */
int agpu_fused_chain_int(agpu_device* dev, int dtype, const void* in, const uint32_t* vin,
                         const agpu_chain_step* steps, int n_steps, void* out, size_t n, uint32_t* vout);

/* ---- routines: crates/routines ---- */

/* Swizzle::merge_op, routines/src/lib.rs:82-120, bool.rs:49-87:
 * out[i] = mask bit i ? a[i] : b[i]; dtype AGPU_BOOL merges bitmaps.
 * validity (merge.rs:17-86): vout = ((va & m) | (vb & ~m)) & vmask with NULL = all ones
 * (SURVEY.md Q7); vout may be NULL when va, vb and vmask are all NULL. */
int agpu_merge(agpu_device* dev, int dtype, const void* a, const void* b, const uint32_t* mask,
               void* out, size_t n, const uint32_t* va, const uint32_t* vb,
               const uint32_t* vmask, uint32_t* vout);

/* Swizzle::take_op, routines/src/lib.rs:122-143, take.rs:9-55, bool.rs:15-46:
 * out[j] = src[idx[j]] for j < m; idx[j] >= src_len reads as zero (robust buffer access).
 * dtype AGPU_BOOL gathers bits.  If vsrc != NULL, vout gets the gathered validity bits. */
int agpu_take(agpu_device* dev, int dtype, const void* src, size_t src_len, const uint32_t* idx,
              void* out, size_t m, const uint32_t* vsrc, uint32_t* vout);

/* Swizzle::put_op, routines/src/lib.rs:145-170, put.rs:9-56, bool/put.wgsl:
 * dst[dst_idx[i]] = src[src_idx[i]] in place, i < m (duplicate dst indices: any one wins).
 * The shader relies on wgpu's robust buffer access; same contract here: src_idx[i] >= src_len
 * reads zero, dst_idx[i] >= dst_len writes nothing (lengths in rows; bits for AGPU_BOOL). */
int agpu_put(agpu_device* dev, int dtype, const void* src, size_t src_len, const uint32_t* src_idx,
             void* dst, size_t dst_len, const uint32_t* dst_idx, size_t m);

/* filter / compaction (named by BASELINE.json config 5; not in the reference — SURVEY a18).
 * Keeps rows whose mask bit is 1 and (if vmask) whose mask is valid, order preserving.
 * Two calls so a sharded caller can learn the count (and exchange it between GPUs) before it
 * allocates the output:
 *   agpu_filter_count  : per-tile selected-row counts and scanned per-group output offsets
 *                        into `scratch` (agpu_filter_scratch_bytes(n) bytes, 16-byte aligned)
 *                        and the total into *total_dev (a device uint64)
 *   agpu_filter_scatter: uses the counts/offsets in `scratch` to compact the values into
 *                        out[0 .. total) and, when vsrc and vout are given, the validity bits
 *                        into vout.  `out_capacity` = rows `out` can hold (and vout:
 *                        ceil(out_capacity/32) words, zeroed here first); rows beyond it are
 *                        dropped, so a caller may launch the scatter BEFORE it knows the total
 *                        (out_capacity = n is always enough) and read the total later. */
size_t agpu_filter_scratch_bytes(size_t n);
int agpu_filter_count(agpu_device* dev, const uint32_t* mask, const uint32_t* vmask, size_t n,
                      void* scratch, uint64_t* total_dev);
int agpu_filter_scatter(agpu_device* dev, int dtype, const void* src, const uint32_t* vsrc,
                        const uint32_t* mask, const uint32_t* vmask, size_t n, void* scratch,
                        void* out, uint32_t* vout, size_t out_capacity);

/* ---- count / offset exchange between the shards of one box, device side (north_star (4)) ----
 * Every rank (one process per GPU) owns a slot area of agpu_exchange_bytes(world) bytes in
 * agpu_ipc_alloc memory, zeroed once, exported with agpu_ipc_export and opened by every peer.
 *   agpu_exchange_post: stores this rank's u64 `*value_dev` (< 2^40) into slot[rank] of EVERY
 *                       rank's area, straight over NVLink peer memory.  peer_slots = HOST array of
 *                       `world` device pointers (entry `rank` = the local area).
 *   agpu_exchange_wait: waits ON THE GPU until all `world` values of exchange `seq` have arrived in
 *                       the local area, then writes out_dev[0..world) = exclusive prefix sums
 *                       (global offset of each rank's output), out_dev[world] = total,
 *                       out_dev[world+1] = status (0 ok, 1 = timed out after timeout_ms; 0 -> 10 s)
 *                       and out_dev[world+2 .. 2*world+2) = the raw per-rank values.
 * Both only enqueue a one-warp kernel on the handle's stream: a sharded filter runs
 * count -> post -> scatter -> wait with no host synchronisation.  `seq` counts the exchanges of
 * this group (0, 1, 2, ...; same on every rank); every rank must post and wait each of them. */
#define AGPU_EXCHANGE_RING 4
size_t agpu_exchange_bytes(int world);
int agpu_exchange_post(agpu_device* dev, const uint64_t* value_dev, void* const* peer_slots, int rank,
                       int world, uint32_t seq);
int agpu_exchange_wait(agpu_device* dev, const void* my_slots, int world, uint32_t seq,
                       uint64_t* out_dev, uint32_t timeout_ms);
/* agpu_filter_count + agpu_exchange_post in one launch: the CTA of the count kernel that finishes
 * last (it scans the group totals) also stores the shard's total into every peer's slot area. */
int agpu_filter_count_post(agpu_device* dev, const uint32_t* mask, const uint32_t* vmask, size_t n,
                           void* scratch, uint64_t* total_dev, void* const* peer_slots, int rank,
                           int world, uint32_t seq);

/* ---- "next" rows ---- */

/* Global-index take across the row-range shards of one column (8f rank 2): every process owns
 * one shard; shards are made visible to the other GPUs of the box through CUDA IPC and the
 * gather kernel reads them directly over NVLink/NVSwitch peer memory — no all-to-all of
 * requests and replies.  Exportable buffers come from cudaMalloc (pool memory cannot be
 * exported). */
int agpu_ipc_alloc(agpu_device* dev, size_t bytes, void** out);
int agpu_ipc_free(agpu_device* dev, void* ptr);
int agpu_ipc_export(agpu_device* dev, const void* ptr, unsigned char handle[AGPU_IPC_HANDLE_BYTES]);
int agpu_ipc_open(agpu_device* dev, const unsigned char handle[AGPU_IPC_HANDLE_BYTES], void** out);
int agpu_ipc_close(agpu_device* dev, void* ptr);
/* out[j] = column[idx[j]] where global row r lives in shard s with shard_begin[s] <= r <
 * shard_begin[s+1] at shard_values[s][r - shard_begin[s]] (local or peer device memory).
 * shard_validity may be NULL (no shard has a bitmap) or hold one bitmap pointer (or NULL = all
 * valid) per shard; vout receives the gathered validity bits.  Rows >= shard_begin[n_shards]
 * read as zero.  All arrays of pointers/offsets are HOST arrays of n_shards (+1) entries. */
int agpu_take_sharded(agpu_device* dev, int dtype, int n_shards, const void* const* shard_values,
                      const uint32_t* const* shard_validity, const uint64_t* shard_begin,
                      const uint32_t* idx, void* out, size_t m, uint32_t* vout);

/* Broadcast::broadcast_op, array/src/kernels/broadcast.rs:6-17, */
/* array/compute_shaders/{f32,i32,u32}/broadcast.wgsl: out[i] = *scalar_host (by value bits) */
int agpu_broadcast(agpu_device* dev, int dtype, const void* scalar_host, void* out, size_t n);

/* Sum::sum_op, arithmetic/src/aggregate_kernels.rs:24-52 + */
/* compute_shaders/{f32,i32,u32}/aggregate.wgsl: same 256-wide pairwise tree order, so the
 * f32 result is bit-identical to the reference's.  out_dev: one element of dtype. */
int agpu_sum(agpu_device* dev, int dtype, const void* a, size_t n, void* out_dev);

/* LogicalContains::{any,all}, logical/src/boolean.rs:106-147: result_dev is a device
 * uint32 (1/0).  `all` counts only the first n_bits. */
int agpu_any(agpu_device* dev, const uint32_t* bits, size_t n_bits, uint32_t* result_dev);
int agpu_all(agpu_device* dev, const uint32_t* bits, size_t n_bits, uint32_t* result_dev);

#ifdef __cplusplus
}
#endif
#endif /* AGPU_H */
