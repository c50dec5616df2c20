/*
 * zdw_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the adobe/zdw hot path (TSV -> ZDW v11 encode and
 * ZDW v9..v11 -> TSV decode).  It exists so that tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg can check the CUDA product bit-for-bit.  Nothing in the product (zdw_b200/)
 * may include, link or call it: the product fails loudly when its CUDA library is missing.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against the
 * reference's own golden vectors (test.zdw v9, analytics-hits.zdw v10, movie_tickets.zdw v10,
 * committed under tests/golden/) and tests/test_oracle_vs_ref.py checks it against the compiled,
 * unmodified reference (oracle/_ref/) on an adversarial corpus.
 *
 * Every function cites the reference file:line (relative to /root/reference/cplusplus) it restates.
 */
#ifndef ZDW_ORACLE_H
#define ZDW_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* column type ids stored on disk: zdw_column_type_constants.h:17-35 */
enum {
  ZO_VARCHAR = 0, ZO_TEXT = 1, ZO_DATETIME = 2, ZO_CHAR_2 = 3, ZO_VISID_LOW = 4, ZO_VISID_HIGH = 5,
  ZO_CHAR = 6, ZO_TINY = 7, ZO_SHORT = 8, ZO_LONG = 9, ZO_LONGLONG = 10, ZO_DECIMAL = 11,
  ZO_TINY_SIGNED = 12, ZO_SHORT_SIGNED = 13, ZO_LONG_SIGNED = 14, ZO_LONGLONG_SIGNED = 15,
  ZO_TINYTEXT = 16, ZO_MEDIUMTEXT = 17, ZO_LONGTEXT = 18
};

/* error codes: ConvertToZDW.h:44-70 (encode) / UnconvertFromZDW.h:34-56 (decode) */
enum {
  ZO_OK = 0,
  ZO_ENC_DESC_FILE_MISSING_TYPE_INFO = 14,
  ZO_ENC_WRONG_NUM_OF_COLUMNS_ON_A_ROW = 15,
  ZO_ENC_OUT_OF_MEMORY = 7,
  ZO_DEC_GZREAD_FAILED = 2,
  ZO_DEC_UNSUPPORTED_ZDW_VERSION_ERR = 5,
  ZO_DEC_ZDW_LONGER_THAN_EXPECTED_ERR = 6,
  ZO_DEC_ROW_COUNT_ERR = 8,
  ZO_DEC_CORRUPTED_DATA_ERROR = 9
};

typedef struct {
  uint32_t ncols;
  char** names;       /* ncols NUL-terminated names */
  uint8_t* types;     /* ncols type ids */
  uint16_t* charsize; /* ncols */
} zo_schema;

/* ConvertToZDW::ReadDescFile, ConvertToZDW.cpp:91-162.  Returns 0 or ZO_ENC_DESC_FILE_MISSING_TYPE_INFO. */
int zo_parse_desc(const char* text, size_t len, zo_schema* out);
void zo_schema_free(zo_schema* s);

typedef struct {
  int trim_trailing_spaces;  /* -t, ConvertToZDW.cpp:295-313 */
  uint32_t rows_per_block;   /* 0 = one block; else close a block every N rows (explicit policy) */
  /* metadata: nmeta pairs, keys already sorted like std::map (ConvertToZDW.cpp:676-694) */
  uint32_t nmeta;
  const char* const* meta_keys;
  const char* const* meta_vals;
  /* explicit block plan (overrides rows_per_block when nplan > 0): block k closes after plan_rows[k] rows; the first
     plan_spill[k] columns of the row that follows have already gone through pass 1 by then - the reference's
     interrupted row (ConvertToZDW.cpp:334-355,404-413; stringheap.cpp:75-86): its strings stay in the
     closed block's dictionary, its numbers in the column ranges, the row itself opens the next block.  Rows left after
     the plan form one last block. */
  uint32_t nplan;
  const uint32_t* plan_rows;
  const uint32_t* plan_spill;
  /* The reference's OWN block cut as a function of the input and one number (used when nplan == 0 and rows_per_block
     == 0; 0 = off).  New strings are packed in first-occurrence order into heap blocks of 64 MiB - a block of
     max(len+1, 64 MiB) is opened whenever the current one has fewer than len+1 bytes free, the rest of the old one is
     wasted (StringHeap::copyToHeap, stringheap.cpp:31-59) - and the insert that opens the heap_blocks-th heap block of
     a ZDW block finds the process over its --mem-limit (allocBlock, stringheap.cpp:75-86): pass 1 stops inside that
     row, after that column (ConvertToZDW.cpp:334-355,404-413).  Which allocation is the first one over the limit
     depends on the virtual size of the reference's process, i.e. on --mem-limit, the allocator and the machine;
     everything else is a pure function of the input.  The same number is applied to every block of
     the file (the 64 MiB blocks are mmap'd, so a finished block gives its memory back). */
  uint32_t heap_blocks;
} zo_encode_opts;

typedef struct {
  uint64_t total_rows;
  uint32_t nblocks;
  uint32_t longest_line;   /* final m_LongestLine */
  uint32_t bad_row;        /* 1-based "Row %u had the problem" when WRONG_NUM_OF_COLUMNS */
  uint64_t dict_entries;   /* of the last block */
} zo_encode_info;

/* ConvertToZDW::processFile (ConvertToZDW.cpp:673-894): whole-file image.  *out is malloc'd. */
int zo_encode_file(const zo_schema* s, const uint8_t* tsv, size_t n, const zo_encode_opts* o,
                   uint8_t** out, size_t* out_len, zo_encode_info* info);

/* The file header alone (version, metadata, names, types, char sizes). ConvertToZDW.cpp:673-737 */
int zo_write_file_header(const zo_schema* s, const zo_encode_opts* o, uint8_t** out, size_t* out_len);

typedef struct {
  /* output column map as in UnconvertFromZDW_Base::outputColumns (UnconvertFromZDW.cpp:1113-1190):
     NULL = all columns in file order; else out_col[c] = output position or -1 (IGNORE). */
  const int* out_col;
  uint32_t n_out;            /* number of output positions when out_col != NULL */
  char sep;                  /* '\t' for files; '\0' for the in-memory API */
} zo_decode_opts;

typedef struct {
  uint16_t version;
  uint32_t ncols;
  uint64_t total_rows;
  uint32_t nblocks;
  uint32_t line_length;      /* exportFileLineLength of the last block */
  size_t consumed;           /* bytes of the image consumed */
  uint32_t block_rows[16];   /* numRows of the first 16 blocks (block-plan recovery in the tests) */
  uint64_t block_offset[16]; /* byte offset of their block headers inside the image */
} zo_decode_info;

/* UnconvertFromZDWToFile<BufferedOutput>::unconvert data path (UnconvertFromZDW.cpp:1030-1219,
 * 758-1000, 1270-1464, 1814-1834).  *out is malloc'd TSV. */
int zo_decode_file(const uint8_t* zdw, size_t n, const zo_decode_opts* o,
                   uint8_t** out, size_t* out_len, zo_decode_info* info);

/* Header parse only: fills a schema (names/types/charsize) and the header length. */
int zo_read_file_header(const uint8_t* zdw, size_t n, zo_schema* s, uint16_t* version, size_t* hdr_len);

void zo_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
