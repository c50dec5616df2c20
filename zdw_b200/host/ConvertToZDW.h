// ConvertToZDW.h -- host side of the TSV -> ZDW encoder, B200 build.
//
// Same public surface as the reference class (cplusplus/ConvertToZDW.h:28-103): constructor flags, `compressor`,
// trimTrailingSpaces, setStatusOutputCallback, loadMetadataFile, convertFile and the ERR_CODE values.  What differs is
// behind it: pass 1, the dictionary, the column statistics and pass 2 of every block run on the GPU through
// zdwb_encode_block (include/zdw_b200.h); this class keeps the .desc.sql rules, the file header, the compressor
// pipe, block stitching, validation and the temp-name + rename protocol.
#ifndef ZDWB_HOST_CONVERTTOZDW_H
#define ZDWB_HOST_CONVERTTOZDW_H

#include <stdint.h>
#include <stdio.h>

#include <map>
#include <string>
#include <vector>

#include "gpu_session.h"
#include "zdw/includes.h"
#include "zdw/status_output.h"

namespace adobe {
namespace zdw {

class ConvertToZDW {
 public:
  static const int CONVERT_ZDW_CURRENT_VERSION;
  static const char CONVERT_ZDW_VERSION_TAIL[3];

  enum Compressor { GZIP = 0, BZIP2 = 1, XZ = 2, FXZ = 3, ZSTD = 4 };

  // values are API (reference ConvertToZDW.h:44-70)
  enum ERR_CODE {
    OK = 0, NO_ARGS = 1, CONVERSION_FAILED = 2, UNTAR_FAILED = 3, MISSING_DESC_FILE = 4, MISSING_SQL_FILE = 5,
    FILE_CREATION_ERR = 6, OUT_OF_MEMORY = 7, UNCONVERT_FAILED = 8, FILE_SIZES_DIFFER = 9, FILES_DIFFER = 10,
    MISSING_ARGUMENT = 11, GZIP_FAILED = 12, BZIP2_FAILED = 13, DESC_FILE_MISSING_TYPE_INFO = 14,
    WRONG_NUM_OF_COLUMNS_ON_A_ROW = 15, BAD_PARAMETER = 16, TOO_MANY_INPUT_FILES = 17, NO_INPUT_FILES = 18,
    CANT_OPEN_TEMP_FILE = 19, UNKNOWN_ERROR = 20, BAD_METADATA_PARAM = 21, BAD_METADATA_FILE = 22,
    ERR_CODE_COUNT
  };
  static const char ERR_CODE_TEXTS[ERR_CODE_COUNT][30];

  ConvertToZDW(const bool bQuiet = false, const bool bStreamingInput = false);
  ~ConvertToZDW();

  void setStatusOutputCallback(StatusOutputCallback cb) { statusOutput = cb; }
  void trimTrailingSpaces(bool val = true) { bTrimTrailingSpaces = val; }
  const char* getInputFileExtension() const { return "sql"; }

  // Returns 0 = success, -1 = cannot open, else the 1-based number of the first line without '='.
  static int loadMetadataFile(const char* filepath, std::map<std::string, std::string>& metadata);

  ERR_CODE convertFile(const char* infile, const char* exeName, const bool bValidate, char* filestub,
                       const char* outputDir = NULL, const char* zArgs = NULL,
                       const std::map<std::string, std::string>& metadata = std::map<std::string, std::string>());

  Compressor compressor;

  // ---- additions of this build (never change the meaning of a reference flag) -------------------------------
  // Block policy.  The reference closes a block when the process runs out of --mem-limit virtual memory
  // (stringheap.cpp:75-86); here a block is closed after `rows` rows (0 = no row limit) or when the rows no
  // longer fit a window of `bytes` TSV bytes, whichever comes first.
  void setRowsPerBlock(uint64_t rows) { rowsPerBlock = rows; }
  void setBlockBytes(size_t bytes) { blockBytes = bytes; }
  // explicit block plan: block k closes after rows[k] rows and the first spill[k] columns of the row that follows count
  // for its dictionary and column ranges (the reference's interrupted row); rows left form a last block
  void setBlockPlan(const std::vector<std::pair<uint64_t, uint32_t> >& plan) { blockPlan = plan; }
  void setGpuDevice(int device) { gpuDevice = device; }
  // Close every block where the reference would when its process is found over --mem-limit at the K-th allocation of a
  // 64 MiB string-heap block (zdwb_encode_opts.heap_blocks).  0 = off.
  void setHeapBlocks(uint32_t k) { heapBlocks = k; }
  // Whole blocks are dealt out to several GPUs: `devices` lists the CUDA device of every encode worker
  // group (a device may appear more than once), `lanes` = workers (context + pinned window + host thread) per entry.
  // The output is the same as with one worker; blocks are written in file order.  Applies to regular input files cut
  // by --block-bytes windows (the default); streamed input, --rows-per-block and --block-plan run on one context.
  void setGpus(const std::vector<int>& devices) { gpuList = devices; }
  void setLanesPerGpu(int lanes) { lanesPerGpu = lanes; }

 private:
  struct DescSchema {
    std::vector<std::string> names;
    std::vector<unsigned char> types;
    std::vector<int> charSizes;
  };
  static bool readDescFile(FILE* f, DescSchema& out);
  static bool metadataIsValid(const std::map<std::string, std::string>& metadata);
  const char* compressorExtension() const;
  const char* compressorCommand() const;

  ERR_CODE processFile(FILE* in, const char* filestub, const DescSchema& schema, const bool bValidate, const char* exeName,
                       const char* outputDir, const char* zArgs, const std::map<std::string, std::string>& metadata);
  ERR_CODE validate(const char* zdwFile, const std::vector<std::string>& srcFiles, const char* exeName,
                    const char* outputDir);
  static void writeFileHeader(FILE* out, const DescSchema& schema, const std::map<std::string, std::string>& metadata);

  StatusOutputCallback statusOutput;
  const bool bQuiet;
  bool bTrimTrailingSpaces;
  const bool bStreamingInput;
  uint64_t rowsPerBlock;
  std::vector<std::pair<uint64_t, uint32_t> > blockPlan;
  size_t blockBytes;
  uint32_t heapBlocks;
  int gpuDevice;
  std::vector<int> gpuList;
  int lanesPerGpu;
  GpuSession gpu;
  struct ParallelOutcome;
  ERR_CODE encodeWindowsParallel(FILE* in, size_t windowBytes, const DescSchema& schema, const char* filestub, const char* exeName,
                                 class AsyncWriter& writer, ParallelOutcome& outcome);
};

}  // namespace zdw
}  // namespace adobe
#endif
