// UnconvertFromZDW.h -- host side of the ZDW -> TSV decoder, B200 build.
//
// Public surface of the reference header (cplusplus/zdw/UnconvertFromZDW.h:34-354): the ERR_CODE and
// COLUMN_INCLUSION_RULE enums, ZDWException, UnconvertFromZDW_Base, UnconvertFromZDWToFile<T> for the two output
// policies, and UnconvertFromZDWToMemory - the row-at-a-time API that test_unconvert_api.cpp drives.  The per-block
// work (block header, dictionary, every row) is one zdwb_decode_block call (include/zdw_b200.h); this layer keeps
// the file header, column selection, .desc.sql / .metadata emitters, status text and the API state machine.
// Files of version 9, 10 and 11 are read (older layouts use the retired tree dictionary and are rejected).
#ifndef ZDWB_HOST_UNCONVERTFROMZDW_H
#define ZDWB_HOST_UNCONVERTFROMZDW_H

#include <stdint.h>
#include <stdio.h>

#include <condition_variable>
#include <map>
#include <mutex>
#include <ostream>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../gpu_session.h"
#include "includes.h"
#include "status_output.h"

namespace adobe {
namespace zdw {

// values are API ("don't change", reference :34-56)
enum ERR_CODE {
  OK = 0, BAD_PARAMETER = 1, GZREAD_FAILED = 2, FILE_CREATION_ERR = 3, FILE_OPEN_ERR = 4,
  UNSUPPORTED_ZDW_VERSION_ERR = 5, ZDW_LONGER_THAN_EXPECTED_ERR = 6, UNEXPECTED_DESC_TYPE = 7, ROW_COUNT_ERR = 8,
  CORRUPTED_DATA_ERROR = 9, HEADER_NOT_READ_YET = 10, HEADER_ALREADY_READ_ERR = 11, AT_END_OF_FILE = 12,
  BAD_REQUESTED_COLUMN = 13, NO_COLUMNS_TO_OUTPUT = 14, PROCESSING_ERROR = 15, UNSUPPORTED_OPERATION = 16,
  METADATA_KEY_NOT_PRESENT = 17,
  ERR_CODE_COUNT
};

enum COLUMN_INCLUSION_RULE {
  FAIL_ON_INVALID_COLUMN,
  SKIP_INVALID_COLUMN,
  EXCLUDE_SPECIFIED_COLUMNS,
  PROVIDE_EMPTY_MISSING_COLUMNS
};

class ZDWException : public std::runtime_error {
 public:
  explicit ZDWException(const ERR_CODE errcode);
  ERR_CODE code;
};

namespace internal {

struct MetadataOptions {
  bool bOutputOnlyMetadata;
  bool bOnlyMetadataKeys;
  bool bAllowMissingKeys;
  std::set<std::string> keys;
  MetadataOptions() : bOutputOnlyMetadata(false), bOnlyMetadataKeys(false), bAllowMissingKeys(false) {}
};

// Byte source: `popen(<decompressor> file)`, the file itself when it is not compressed, or stdin; buffered ahead (plain
// memory that grows as the bytes arrive) so that a whole block can be handed to the GPU in one piece.
class ZdwInput {
 public:
  ZdwInput();
  ~ZdwInput();
  bool openCommand(const std::string& cmd);
  bool openFile(const std::string& path);  // an uncompressed file is read directly (the reference pipes it through cat)
  void openStdin();
  bool is_open() const { return fp != NULL; }
  // makes at least n bytes available at data() unless the stream ends first; returns what is available
  size_t ensure(size_t n);
  const char* data() const { return buf + pos; }
  size_t available() const { return len - pos; }
  void consume(size_t n);
  bool sourceEnded() const { return ended; }
  // like BufferedInput::eof(): true once a read has run into the end of the stream with nothing buffered
  bool eof() const { return eofSeen; }
  void noteEofProbe();
  void finalDummyRead();
  unsigned long long offset() const { return consumedTotal; }

 private:
  ZdwInput(const ZdwInput&);
  ZdwInput& operator=(const ZdwInput&);
  FILE* fp;
  bool isPipe;
  char* buf;
  size_t cap, len, pos;
  bool ended, eofSeen;
  unsigned long long consumedTotal;
};

}  // namespace internal

// Output policies of the reference (BufferedOutput.h:23-196).  Column selection and ordering happen on the GPU, so
// both are thin FILE* writers here; the two names are kept because callers spell them as template arguments.
// writeLater() hands a whole block of rows to a writer thread: the caller goes on to read and decode the next block
// while this one is written (one block in flight; every call on the object first waits for it).
class BufferedOutput {
 public:
  explicit BufferedOutput(FILE* f) : fp(f), jobData(NULL), jobSize(0), busy(false), stop(false), started(false), failed(false) {}
  ~BufferedOutput();
  bool write(const void* data, size_t size) {
    waitIdle();
    return !size || fwrite(data, 1, size, fp) == size;
  }
  // `data` must stay untouched until the next call on this object returns
  void writeLater(const void* data, size_t size);
  bool waitIdle();  // false once a deferred write came up short
  FILE* file() const { return fp; }

 private:
  BufferedOutput(const BufferedOutput&);
  BufferedOutput& operator=(const BufferedOutput&);
  void run();
  FILE* fp;
  std::mutex m;
  std::condition_variable cv;
  const void* jobData;
  size_t jobSize;
  bool busy, stop, started, failed;
  std::thread worker;
};
class BufferedOrderedOutput : public BufferedOutput {
 public:
  explicit BufferedOrderedOutput(FILE* f) : BufferedOutput(f) {}
};

class UnconvertFromZDW_Base {
 public:
  static const int UNCONVERT_ZDW_VERSION;
  static const char UNCONVERT_ZDW_VERSION_TAIL[3];
  static const char ERR_CODE_TEXTS[ERR_CODE_COUNT + 1][30];

  UnconvertFromZDW_Base(const std::string& inFileName, const bool bShowStatus = true, const bool bQuiet = true,
                        const bool bTestOnly = false, const bool bOutputDescFileOnly = false);
  virtual ~UnconvertFromZDW_Base();

  void setStatusOutputCallback(StatusOutputCallback cb) { statusOutput = cb; }
  static std::string getVersion();

  std::vector<std::string> getColumnNames() const { return columnNames; }
  UCHAR* getColumnTypes() const { return const_cast<UCHAR*>(columnType.data()); }
  ULONG getRowsRead() const { return rowsRead; }   // in the current block
  ULONG getNumLines() const { return numLines; }   // in the current block
  bool isLastBlock() const { return lastBlock != 0; }
  bool isFinished() const { return input && input->eof(); }
  bool isReadOpen() const { return input && input->is_open(); }

  void printError(const std::string& exeName, const std::string& inFileName);

  void outputNonEmptyColumnHeader(bool bFlag = true) { bOutputNonEmptyColumnHeader = bFlag; }
  ERR_CODE readHeader();
  bool setNamesOfColumnsToOutput(const std::string& csv_str, COLUMN_INCLUSION_RULE inclusionRule);
  bool setNamesOfColumnsToOutput(const std::vector<std::string>& csv_vector, COLUMN_INCLUSION_RULE inclusionRule);
  void showBasicStatisticsOnly(bool bVal = true) { bShowBasicStatisticsOnly = bVal; }
  ERR_CODE GetSchema(std::ostream& stream);
  void setMetadataOptions(const internal::MetadataOptions& options) { metadataOptions = options; }

  void setGpuDevice(int device) { gpuDevice = device; }  // addition of this build
  // Whole blocks are decoded by several workers (a device may be listed more than once; `lanes` workers per entry):
  // the calling thread skims every block for its length (the file has none, reference :782-810,1577-1589), the
  // workers decode, the rows leave in file order.  File / stdout output only; the output does not depend on it.
  void setGpus(const std::vector<int>& devices) { gpuList = devices; }
  void setLanesPerGpu(int lanes) { lanesPerGpu = lanes; }

 protected:
  enum { IGNORE_COLUMN = -1, USE_VIRTUAL_COLUMN = -2 };
  enum STATE { ZDW_BEGIN, ZDW_PARSE_BLOCK_HEADER, ZDW_OUTPUT_BLOCK_HEADER, ZDW_GET_NEXT_ROW, ZDW_FINISHING, ZDW_END };

  struct BlockInfo {      // what the host needs to know about the block in front of the input cursor
    ULONG numLines, lineLength;
    UCHAR last;
    ULONGLONG dictionarySize;
    std::vector<UCHAR> columnSize;
    size_t rowsOffset;    // bytes from the block start to its first row
    size_t maxRowBytes;   // flag bytes + every value present
    size_t numSetColumns; // flag bytes per row
  };
  // Reads the block header that starts at the cursor (readLineLength / readDictionary sizes / readColumnFieldStats,
  // reference UnconvertFromZDW.cpp:758-1000) without consuming it, buffers the whole block and prints the
  // status lines of those functions.
  ERR_CODE peekBlock(BlockInfo& info);
  // Decodes that block on the GPU.  separator '\t' (files) or '\0' (in-memory rows).
  ERR_CODE decodeBlock(const BlockInfo& info, unsigned char separator, bool wantRowOffsets, bool validateOnly,
                       bool wantFlagCounts, zdwb_rows_out* out, GpuSession* session = NULL, bool skimOnly = false);
  // the same on explicit bytes (a complete block when atEnd): what the decode workers call.  Reads only members that
  // do not change after readHeader(); error texts go to *errText instead of the status callback.
  int decodeBytes(GpuSession& g, const void* data, size_t avail, bool atEnd, unsigned long long firstRow, unsigned char separator,
                  bool wantRowOffsets, bool validateOnly, bool wantFlagCounts, bool skimOnly, zdwb_rows_out* out,
                  bool outputOnDevice = false) const;
  std::string getBlockHeaderString(const BlockInfo& info) const;

  ERR_CODE outputDescToFile(const std::vector<std::string>& names, const std::string& outputDir, const char* filestub,
                            const char* ext);
  ERR_CODE outputDescToStdOut(const std::vector<std::string>& names);
  ERR_CODE outputMetadataToFile(const std::string& outputDir, const char* filestub) const;
  ERR_CODE outputMetadataToStdOut() const;
  size_t readBytes(void* buf, const size_t len, const bool bHaltOnReadError = true);

  static std::string GetBaseNameForInFile(const std::string& inFileName);
  static void splitDirAndBase(const std::string& inFileName, std::string& dir, std::string& base);
  bool UseVirtualExportBaseNameColumn() const { return indexForVirtualBaseNameColumn != IGNORE_COLUMN; }
  bool UseVirtualExportRowColumn() const { return indexForVirtualRowColumn != IGNORE_COLUMN; }
  size_t numOutputColumns() const;
  void setState(STATE s) { eState = s; }

  ULONG exportFileLineLength;
  ULONG virtualLineLength;
  std::map<std::string, std::string> metadata;  // version 11+
  USHORT version;
  ULONG numLines;
  ULONG numColumnsInExportFile;
  ULONG numColumns;
  UCHAR lastBlock;
  std::string exeName;
  const std::string inFileName;
  const std::string inFileBaseName;
  internal::ZdwInput* input;

  const bool bOutputDescFileOnly;
  const bool bShowStatus, bQuiet;
  const bool bTestOnly;
  bool bOutputNonEmptyColumnHeader;
  bool bShowBasicStatisticsOnly;
  bool bFailOnInvalidColumns;
  std::map<std::string, unsigned> namesOfColumnsToOutput;
  bool bExcludeSpecifiedColumns;
  bool bOutputEmptyMissingColumns;

  internal::MetadataOptions metadataOptions;
  int indexForVirtualBaseNameColumn;
  int indexForVirtualRowColumn;
  std::vector<std::string> columnNames;
  std::vector<UCHAR> columnType;
  std::vector<USHORT> columnCharSize;   // empty before version 7
  std::vector<int> outputColumns;        // file column -> output position, IGNORE_COLUMN = dropped
  std::map<int, std::string> blankColumnNames;

  ULONG rowsRead;
  unsigned long long rowsBeforeBlock;    // rows of earlier blocks: virtual_export_row runs on across blocks
  StatusOutputCallback statusOutput;
  STATE eState;

  int gpuDevice;
  std::vector<int> gpuList;
  int lanesPerGpu;
  size_t lastBlockBytes;      // size of the previous block: how much of the next one is buffered before the first try
  GpuSession gpu;
  GpuSession gpu2;            // file output only: blocks alternate between two contexts (see parseNextBlock)
  unsigned blocksToSink;      // blocks decoded for a file sink so far

 private:
  std::vector<std::string> getDesc(const std::vector<std::string>& names, const std::string& nameTypeSeparator,
                                   const std::string& delimiter) const;
  std::string getColumnDesc(const std::string& name, UCHAR type, size_t index, const std::string& nameTypeSeparator,
                            const std::string& delimiter) const;
  ERR_CODE outputDesc(const std::vector<std::string>& names, FILE* out);
  ERR_CODE outputMetadata(FILE* out) const;
};

template <typename T>
class UnconvertFromZDW : public UnconvertFromZDW_Base {
 public:
  UnconvertFromZDW(const std::string& inFileName, const bool bShowStatus = true, const bool bQuiet = true,
                   const bool bTestOnly = false, const bool bOutputDescFileOnly = false)
      : UnconvertFromZDW_Base(inFileName, bShowStatus, bQuiet, bTestOnly, bOutputDescFileOnly) {}

 protected:
  ERR_CODE parseNextBlock(T& buffer);
  // every remaining block, decoded by several workers (see setGpus); file-like sinks only
  ERR_CODE decodeBlocksFanOut(T& sink);
};

template <typename BufferedOutput_T>
class UnconvertFromZDWToFile : public UnconvertFromZDW<BufferedOutput_T> {
 public:
  UnconvertFromZDWToFile(const std::string& inFileName, const bool bShowStatus = true, const bool bQuiet = true,
                         const bool bTestOnly = false, const bool bOutputDescFileOnly = false)
      : UnconvertFromZDW<BufferedOutput_T>(inFileName, bShowStatus, bQuiet, bTestOnly, bOutputDescFileOnly), out(NULL) {}

  ERR_CODE unconvert(const char* exeName, const char* outputBasename, const char* ext, const char* outputDir, bool bStdout);

 private:
  FILE* out;
};

// Sink handed to UnconvertFromZDW<T> by the in-memory API (the reference's BufferedOutputInMem,
// BufferedOutput.cpp:304-426): rows are NUL-separated fields, handed out one per getRow call.
class BufferedOutputInMem;

class UnconvertFromZDWToMemory : public UnconvertFromZDW<BufferedOutputInMem> {
 public:
  // With bUseInternalBuffer = false use getRow(char** buffer, size_t* size, ...): the row is copied into the caller's
  // buffer, which is replaced (delete[] / new[]) when it is too small, exactly like the reference.
  UnconvertFromZDWToMemory(const std::string& inFileName, const bool bUseInternalBuffer = true, const bool bShowStatus = true,
                           const bool bQuiet = true, const bool bTestOnly = false, const bool bOutputDescFileOnly = false);
  ~UnconvertFromZDWToMemory();

  ERR_CODE getRow(const char** outColumns);
  ERR_CODE getRow(char** buffer, size_t* size, const char** outColumns, size_t& numColumns);
  ERR_CODE getNumOutputColumns(size_t& num);
  // While the caller walks the rows of a block, the next block is read and decoded on a helper thread in a second GPU
  // context (on by default; off: one thread, one context, like the reference's one-thread loop).  Call before the
  // first getRow.
  void setDecodeAhead(bool on) { bDecodeAhead = on; }
  size_t getCurrentRowLength();
  // valid after getNumOutputColumns or getRow
  ULONG getLineLength() { return this->exportFileLineLength + this->virtualLineLength; }
  void getColumnNamesVector(std::vector<std::string>& columnNamesVector);
  bool hasColumnName(const std::string& name) const;
  bool OutputDescToFile(const std::string& outputDir);
  std::vector<std::pair<uint64_t, std::string> > getFileLineage();

 protected:
  ERR_CODE handleZDWParseBlockHeader();

 private:
  ERR_CODE deliver(const char* src, size_t len, size_t fields, char** buffer, size_t* size, const char** outColumns);
  bool bUseInternalBuffer;
  bool blockOpen;            // a decoded block is waiting to be handed out
  size_t neededBufferSize;   // line length + virtual + 1, or the block header line
  std::string pendingHeaderLine;
  const char* slab;          // decoded rows of the current block (owned by the GPU context)
  const uint64_t* slabRowOff;
  size_t currentRowLength;
  std::vector<char> internalRow;
  // decode-ahead (see setDecodeAhead): the block behind the current one, decoded in the context the current block does
  // not live in.  The helper only runs between two handleZDWParseBlockHeader calls; everything that touches `input`
  // joins it first.
  void startDecodeAhead();
  void joinDecodeAhead();
  bool bDecodeAhead;
  int slabSession;           // 0: the current block's rows live in `gpu`, 1: in `gpu2`
  std::thread aheadThread;
  bool aheadPending;
  int aheadRc;               // ZDWB_* of the helper's decode (-1: it did not get that far)
  zdwb_rows_out aheadRows;
};

}  // namespace zdw
}  // namespace adobe
#endif
