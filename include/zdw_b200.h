/*
 * zdw_b200.h -- C ABI of libzdw_b200.so, the B200 (sm_100a) implementation of the adobe/zdw hot path.
 *
 * The reference (adobe/zdw, cplusplus/) has no plugin/FFI interface: its hot path sits inside two
 * C++ classes.  This ABI is the seam a maintainer cuts at: the bodies of the functions cited below
 * are replaced by one call each, everything around them (CLI, .desc.sql, file header, compressor
 * pipe, block stitching) stays host C++.  See INTEGRATION.md for the binding on the reference side.
 *
 *   zdwb_encode_block   replaces, for one ZDW block,
 *        ConvertToZDW::parseInput            ConvertToZDW.cpp:329-414   (pass 1)
 *        GetNextRow / GetDataRow             getnextrow.cpp:26-84, ConvertToZDW.cpp:265-323, :1048-1067
 *        Dictionary::insert/write/getOffset  dictionary.cpp:31-111
 *        writeLookupColumnStats              ConvertToZDW.cpp:417-483
 *        writeBlockRows                      ConvertToZDW.cpp:486-606   (pass 2)
 *        and the block header fields         ConvertToZDW.cpp:839-842
 *   zdwb_decode_block   replaces, for one ZDW block,
 *        UnconvertFromZDW_Base::parseBlockHeader   UnconvertFromZDW.cpp:758-1000
 *        UnconvertFromZDW<T>::readNextRow          UnconvertFromZDW.cpp:1270-1464 (+GetWord :359-371,
 *        llutoa/lltoa :318-356, outputDefault :1224-1266) for all rows of the block
 *
 * Plain C types only; int status codes; no exceptions and no C++ types cross the boundary.  Outputs
 * are owned by the context and stay valid until the next call on the same context.  One context per
 * host thread / GPU.  There is no CPU fallback: every entry point fails with ZDWB_ERR_NO_DEVICE or
 * ZDWB_ERR_CUDA when the GPU path is unavailable.
 */
#ifndef ZDW_B200_H
#define ZDW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZDWB_ABI_VERSION 4

/* status codes (mapped to the reference's ERR_CODE enums by the host classes) */
enum {
  ZDWB_OK = 0,
  ZDWB_ERR_CUDA = 1,          /* CUDA runtime/driver error                     -> PROCESSING_ERROR / CONVERSION_FAILED */
  ZDWB_ERR_OOM = 2,           /* device or pinned allocation failed            -> OUT_OF_MEMORY (ConvertToZDW.h:52) */
  ZDWB_ERR_WRONG_COLUMNS = 3, /* a row's field count != schema                 -> WRONG_NUM_OF_COLUMNS_ON_A_ROW (15) */
  ZDWB_ERR_BAD_ARG = 4,       /* NULL / inconsistent arguments                 -> BAD_PARAMETER */
  ZDWB_ERR_UNSUPPORTED = 5,   /* input outside the supported envelope (see DESIGN.md limits) */
  ZDWB_ERR_CORRUPT = 6,       /* dictionary offset out of range                -> CORRUPTED_DATA_ERROR (9) */
  ZDWB_ERR_TRUNCATED = 7,     /* block runs past the bytes supplied            -> GZREAD_FAILED (2) */
  ZDWB_ERR_ROW_COUNT = 8,     /* fewer rows than the block header promises     -> ROW_COUNT_ERR (8) */
  ZDWB_ERR_NO_DEVICE = 9      /* no usable CUDA device */
};

/* column type ids as stored in the file: zdw_column_type_constants.h:17-35 */
enum {
  ZDWB_VARCHAR = 0, ZDWB_TEXT = 1, ZDWB_DATETIME = 2, ZDWB_CHAR_2 = 3, ZDWB_VISID_LOW = 4,
  ZDWB_VISID_HIGH = 5, ZDWB_CHAR = 6, ZDWB_TINY = 7, ZDWB_SHORT = 8, ZDWB_LONG = 9, ZDWB_LONGLONG = 10,
  ZDWB_DECIMAL = 11, ZDWB_TINY_SIGNED = 12, ZDWB_SHORT_SIGNED = 13, ZDWB_LONG_SIGNED = 14,
  ZDWB_LONGLONG_SIGNED = 15, ZDWB_TINYTEXT = 16, ZDWB_MEDIUMTEXT = 17, ZDWB_LONGTEXT = 18
};

typedef struct zdwb_ctx zdwb_ctx;

/* ---- context -------------------------------------------------------------------------------- */

/* Creates a context bound to CUDA device `device`.  `workspace_hint` (bytes, 0 = default) seeds the
 * device memory pool so that steady-state calls do not allocate. */
int zdwb_ctx_create(int device, size_t workspace_hint, zdwb_ctx** out);
void zdwb_ctx_destroy(zdwb_ctx* ctx);

/* Human-readable description of the last failure on this context ("" if none). */
const char* zdwb_last_error(const zdwb_ctx* ctx);

/* Run all work of this context on the given cudaStream_t (NULL = the context's own stream).  Lets a
 * harness time the kernels with events recorded on its own stream. */
int zdwb_ctx_set_stream(zdwb_ctx* ctx, void* cuda_stream);

/* Tuning / test knobs (name = value); every default is what the measurements in DESIGN.md picked.  Encode:
 * "enc_delta" (pass-1 variant: 1 = row-delta, 0 = general, -1 = by row width), "enc_dtile" (bytes per row-delta tile, a
 * power of two >= 2048; 0 = by input size), "enc_p2_rows" (rows per pass-2 tile, 0 = automatic), "small_sort_max"
 * (largest dictionary ranked by the tile-sort path instead of the radix sort), "sort_radix_items" (records per thread
 * in the radix passes: 4 or 16, 0 = by size), "ht_initial_log2" (first-try size of the string hash set).  Decode:
 * "dec_delta" (0/1: rows of wide schemas are assembled from the row before), "dec_strip_rows" (rows per strip of the row
 * kernels, 0 = automatic), "dec_group_lanes" (lanes per row in the narrow-schema kernels: 8, 16, 32, 0 = by schema
 * width), "dec_emit_words" (0/1: cached texts leave as aligned words), "dec_tile_bytes" (row-stream bytes per CTA in the
 * row-boundary discovery), "dec_readback_kernel" (0/1: small results reach the host through a kernel's stores into
 * mapped memory instead of the copy engine).  Both: "copy_gate" (0/1, default 1: host<->device copies of 8 MiB and more
 * take turns with those of the process's other contexts on the same device instead of sharing the link),
 * "kernel_timing" (0/1).  Returns ZDWB_ERR_BAD_ARG for unknown names. */
int zdwb_ctx_set_tuning(zdwb_ctx* ctx, const char* name, long long value);

/* Number of kernels this context has launched so far (bench.py reports the delta as gpu_launches). */
unsigned long long zdwb_ctx_kernel_launches(const zdwb_ctx* ctx);

/* With the tuning knob "kernel_timing" = 1 every kernel launch is bracketed by CUDA events on the launching
 * stream.  This call synchronises, writes one line "name<TAB>launches<TAB>total_ms" per kernel into buf
 * (NUL-terminated, truncated to cap), clears the record and returns the untruncated length. */
size_t zdwb_ctx_kernel_times(zdwb_ctx* ctx, char* buf, size_t cap);

int zdwb_abi_version(void);

/* Number of CUDA devices this process can use (0 = none; the host tools spread whole ZDW blocks over them). */
int zdwb_device_count(void);

/* ---- schema ---------------------------------------------------------------------------------- */

typedef struct {
  uint32_t ncols;
  const uint8_t* types; /* ncols type ids (host memory), as produced by ReadDescFile (ConvertToZDW.cpp:91-162) */
} zdwb_schema;

/* ---- encode ---------------------------------------------------------------------------------- */

typedef struct {
  int32_t trim_trailing_spaces; /* -t : ConvertToZDW.cpp:295-313 */
  int32_t input_on_device;      /* `tsv` is a device pointer (already resident in HBM) */
  int32_t output_on_device;     /* leave the encoded block in HBM: out->bytes is then a device pointer */
  int32_t more_input_follows;   /* the buffer is a window of a larger input: an unterminated last line is NOT the dropped
                                   final line of getnextrow.cpp:67-69 but the head of a row that continues in the next
                                   window; it is left unconsumed (tsv_consumed points at it) and isLast is written as 0 */
  uint32_t prev_longest_line;   /* m_LongestLine carried in from earlier blocks of the file (0 = 16384 start value,
                                   ConvertToZDW.cpp:965; the field is cumulative, getnextrow.cpp:57-65) */
  uint32_t spill_cols;          /* with max_rows: the first spill_cols columns of the row AFTER the block's last row also go
                                   through pass 1 - their strings enter this block's dictionary, their numbers its column
                                   ranges - and that row's length counts for longestLine; the row itself is not part of
                                   the block.  This is the reference's interrupted row when it runs low on memory
                                   (ConvertToZDW.cpp:334-355,404-413; stringheap.cpp:75-86): with
                                   (max_rows, spill_cols) per block any file the reference cut on its own can be
                                   reproduced byte for byte.  0 = none. */
  uint64_t max_rows;            /* 0 = every row in the buffer; else close the block after this many rows */
  uint32_t heap_blocks;         /* K > 0 (and max_rows == 0): close the block where the REFERENCE would when its process is found
                                   over --mem-limit at the K-th allocation of a 64 MiB string-heap block: new strings are packed
                                   in first-occurrence order like StringHeap::copyToHeap (stringheap.cpp:31-59,75-86), the block
                                   ends in front of the row - and behind the column - whose insert opens heap block K
                                   (ConvertToZDW.cpp:334-355,404-413).  K is the one number of that cut which
                                   depends on the reference's process instead of the input.  If the buffer ends before heap
                                   block K opens and more input follows, nothing is encoded (nrows = 0, tsv_consumed = 0): the
                                   caller widens the window.  K = 1 is the reference's OUT_OF_MEMORY (ZDWB_ERR_OOM). */
  uint32_t reserved_e;
} zdwb_encode_opts;

typedef struct {
  const uint8_t* bytes;   /* the block: numRows u32 | longestLine u32 | isLast u8 | dictionary | columnSize[nc] |
                             columnBase u64[#used] | rows   (SURVEY Appendix A).  isLast is 1 iff the call consumed
                             every row of the buffer; the host stitcher may patch it (offset 8). */
  size_t len;
  uint32_t nrows;         /* rows encoded into this block */
  uint32_t longest_line;  /* value written to the block header (cumulative with prev_longest_line) */
  uint64_t tsv_consumed;  /* bytes of `tsv` covered by this block (start of the next block's first row) */
  uint64_t rows_in_buffer;/* total logical rows found in the buffer */
  uint32_t bad_row;       /* ZDWB_ERR_WRONG_COLUMNS: 1-based row number ("Row %u had the problem") */
  uint32_t ncols_used;    /* columns with columnSize != 0 */
  uint64_t dict_entries;  /* unique strings in the block dictionary (Dictionary::getNumEntries) */
  uint64_t dict_bytes;    /* Dictionary::getSize(): 1 + sum(len+1) */
  uint32_t dict_index_size; /* Dictionary::getBytesInOffset() */
  uint32_t reserved;
} zdwb_block_out;

/* Encodes rows of the TSV buffer `tsv[0..n)` (which must start at a row boundary) into one ZDW block.
 * An input with zero rows yields ZDWB_OK with out->len == 0 and out->nrows == 0 ("Empty data file",
 * ConvertToZDW.cpp:824-835). */
int zdwb_encode_block(zdwb_ctx* ctx, const zdwb_schema* schema, const void* tsv, size_t n,
                      const zdwb_encode_opts* opts, zdwb_block_out* out);

/* ---- decode ---------------------------------------------------------------------------------- */

/* constant text for an output position (virtual_export_basename, UnconvertFromZDW.cpp:1253-1255) */
typedef struct {
  uint32_t pos;              /* output position */
  uint32_t len;
  const char* text;          /* host memory, no separators inside */
} zdwb_fill;

typedef struct {
  int32_t input_on_device;   /* `zdw` is a device pointer */
  int32_t output_on_device;  /* leave the TSV (and row offsets) in HBM */
  int32_t want_row_offsets;  /* also return out->row_off[nrows+1] (needed by the row-at-a-time getRow API) */
  int32_t at_end_of_file;    /* `avail` reaches the end of the file (mirrors input->eof(), UnconvertFromZDW.cpp:1577) */
  uint8_t separator;         /* '\t' (files, BufferedOutput) or '\0' (BufferedOutputInMem); the row terminator is
                                '\n' resp. '\0' */
  uint8_t reserved[7];
  /* Column projection (UnconvertFromZDW_Base::outputColumns, UnconvertFromZDW.cpp:1113-1190).  NULL = every column
   * in file order.  Otherwise out_col[c] is the output position of file column c or -1 to drop it, and n_out is the
   * number of output positions (positions no file column maps to are written as empty fields, the
   * PROVIDE_EMPTY_MISSING_COLUMNS case). */
  const int32_t* out_col;
  uint32_t n_out;
  uint32_t n_fills;          /* constant texts for positions no file column maps to */
  const zdwb_fill* fills;
  int32_t rownum_pos;        /* output position of the running row number (virtual_export_row, :1256-1261), -1 = none;
                                only honoured when out_col != NULL */
  int32_t validate_only;     /* -t: walk every row, bounds-check dictionary offsets, produce no text (:1488-1572) */
  uint64_t first_row_number; /* number of the block's first row (1-based, runs on across blocks) */
  int32_t want_flag_counts;  /* -s: also return out->flag_counts[ncols_used] = rows in which the column's bit is set */
  int32_t skim_only;         /* only find where the block ends: out->consumed (and the header fields) are filled in, no row is
                                formatted.  The file has no block length (UnconvertFromZDW.cpp:782-810,1577-1589): a host that
                                spreads blocks over several GPUs skims block k to learn where block k+1 starts */
} zdwb_decode_opts;

typedef struct {
  const uint8_t* tsv;       /* decoded rows, fields separated by `separator` */
  size_t len;
  const uint64_t* row_off;  /* nrows+1 offsets into tsv (only if want_row_offsets) */
  uint32_t nrows;           /* numRows of the block header */
  uint32_t line_length;     /* exportFileLineLength of the block header */
  uint8_t is_last;          /* isLastBlock */
  uint8_t reserved[7];
  uint64_t consumed;        /* bytes of `zdw` this block occupied: the next block starts at zdw + consumed */
  uint64_t dict_bytes;
  uint32_t ncols_used;
  uint32_t reserved2;
  const uint64_t* flag_counts; /* host memory, ncols_used entries (only if want_flag_counts) */
} zdwb_rows_out;

/* Decodes the block that starts at zdw[0] (the numRows field) given `avail` bytes. `version` is the file's
 * version word (9, 10 or 11: the block layout is identical). */
int zdwb_decode_block(zdwb_ctx* ctx, const zdwb_schema* schema, const void* zdw, size_t avail,
                      const zdwb_decode_opts* opts, zdwb_rows_out* out);

/* ---- pinned host staging (used by the host classes and the end-to-end bench) ------------------ */
void* zdwb_host_alloc(size_t bytes);   /* cudaHostAlloc'd (pinned, usable from every device) memory, NULL on failure */
void zdwb_host_free(void* p);

/* ---- file descriptor <-> device (ABI 4) ---------------------------------------------------------
 * What the host tools use instead of window-sized host buffers: the reference reads its input with fgets
 * (getnextrow.cpp:26-84) and writes rows through BufferedOutput (BufferedOutput.cpp:239-260); here a window of the
 * input file goes straight to the device and the decoded rows of a block straight to the output file, in 8 MiB chunks
 * through a pinned ring owned by the context, the copy of one chunk under the read / write of the next.
 *   zdwb_fd_to_device: `len` bytes of fd from `offset` (pread; offset < 0: read() at the descriptor's position) into a
 *     device buffer of the context; *dev stays valid until the next zdwb_fd_to_device call on it.  Encode it with
 *     zdwb_encode_opts.input_on_device = 1.
 *   zdwb_device_to_fd: `len` device bytes (e.g. zdwb_rows_out.tsv of a call with output_on_device = 1) to fd at `offset`
 *     (pwrite; offset < 0: write()).  Returns when everything is written.
 * ZDWB_ERR_BAD_ARG with zdwb_last_error() set when the descriptor fails or ends early. */
int zdwb_fd_to_device(zdwb_ctx* ctx, int fd, long long offset, size_t len, const void** dev);
int zdwb_device_to_fd(zdwb_ctx* ctx, const void* dev, size_t len, int fd, long long offset);

#ifdef __cplusplus
}
#endif
#endif /* ZDW_B200_H */
