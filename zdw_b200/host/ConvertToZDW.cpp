// ConvertToZDW.cpp -- see ConvertToZDW.h.  Reference behaviour cited as cplusplus/ConvertToZDW.cpp:<line>.
#include "ConvertToZDW.h"

#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <fstream>
#include <new>

#include "block_pipeline.h"
#include "stream_pipeline.h"

#include <atomic>
#include <thread>

using std::map;
using std::string;
using std::vector;

namespace adobe {
namespace zdw {

const int ConvertToZDW::CONVERT_ZDW_CURRENT_VERSION = 11;
const char ConvertToZDW::CONVERT_ZDW_VERSION_TAIL[3] = "b";

const char ConvertToZDW::ERR_CODE_TEXTS[ERR_CODE_COUNT][30] = {
  "OK", "NO_ARGS", "CONVERSION_FAILED", "UNTAR_FAILED", "MISSING_DESC_FILE", "MISSING_SQL_FILE", "FILE_CREATION_ERR",
  "OUT_OF_MEMORY", "UNCONVERT_FAILED", "FILE_SIZES_DIFFER", "FILES_DIFFER", "MISSING_ARGUMENT", "GZIP_FAILED",
  "BZIP2_FAILED", "DESC_FILE_MISSING_TYPE_INFO", "WRONG_NUM_OF_COLUMNS_ON_A_ROW", "BAD_PARAMETER",
  "TOO_MANY_INPUT_FILES", "NO_INPUT_FILES", "CANT_OPEN_TEMP_FILE", "Unknown error", "BAD_METADATA_PARAMETER",
  "BAD_METADATA_FILE"};

namespace {

const size_t DEFAULT_BLOCK_BYTES = (size_t)1 << 30;   // TSV bytes per block window
const size_t MAX_WINDOW_BYTES = 0xff000000ull;         // zdwb_encode_block takes < 4 GiB per call

bool startsWith(const char* s, const char* prefix) { return strncmp(s, prefix, strlen(prefix)) == 0; }

// ZDW_HOST_TIMING=1: wall-clock of the host-side stages on stderr (diagnostics)
bool hostTiming() {
  static const bool on = getenv("ZDW_HOST_TIMING") != NULL;
  return on;
}
double nowSeconds() {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

// the window's buffers: plain memory.  Pinning a window-sized buffer costs about as much as copying out of it unpinned
// once did (0.45 s per GiB) and needs the CUDA context; plain memory can be filled while CUDA is still starting up, and the
// library moves it to the device through its own small pinned ring
void* plainAlloc(size_t n) { return malloc(n); }
void plainFree(void* p) { free(p); }

// What the reference keeps of a streamed input for validation (GetDataRow, ConvertToZDW.cpp:274-283,316-319): the
// rows as GetNextRow returns them - blank lines skipped, an unterminated last line dropped (getnextrow.cpp:39-43,67-69)
// - and with -t every field without its trailing spaces (dump_trimmed_row_to_temp_file, :1069-1090).  p[0..n) are the
// bytes one block covered; it starts at a row boundary.
bool writeRowsForValidation(FILE* to, const char* p, size_t n, bool trim) {
  size_t at = 0;
  while (at < n) {
    // end of the logical line: a newline behind an even number of backslashes
    size_t e = at;
    bool found = false;
    while (e < n) {
      const void* hit = memchr(p + e, '\n', n - e);
      if (!hit) break;
      e = (size_t)(static_cast<const char*>(hit) - p);
      size_t k = e, slashes = 0;
      while (k > at && p[k - 1] == '\\') {
        --k;
        ++slashes;
      }
      if ((slashes & 1u) == 0) {
        found = true;
        break;
      }
      ++e;
    }
    if (!found) break;       // an unterminated last line is not a row
    if (e == at) {           // blank line
      ++at;
      continue;
    }
    if (!trim) {
      if (fwrite(p + at, 1, e - at + 1, to) != e - at + 1) return false;
    } else {
      size_t f = at;
      for (;;) {
        // end of the field: a tab behind an even number of backslashes (get_next_column, :1048-1067)
        size_t t = f;
        while (t < e) {
          if (p[t] == '\t') {
            size_t k = t, slashes = 0;
            while (k > f && p[k - 1] == '\\') {
              --k;
              ++slashes;
            }
            if ((slashes & 1u) == 0) break;
          }
          ++t;
        }
        size_t end = t;
        while (end > f && p[end - 1] == ' ') --end;
        if (end > f && fwrite(p + f, 1, end - f, to) != end - f) return false;
        if (fputc(t < e ? '\t' : '\n', to) == EOF) return false;
        if (t >= e) break;
        f = t + 1;
      }
    }
    at = e + 1;
  }
  return true;
}

}  // namespace

ConvertToZDW::ConvertToZDW(const bool quiet, const bool streamingInput)
    : compressor(GZIP), statusOutput(defaultStatusOutputCallback), bQuiet(quiet), bTrimTrailingSpaces(false),
      bStreamingInput(streamingInput), rowsPerBlock(0), blockBytes(DEFAULT_BLOCK_BYTES), heapBlocks(0), gpuDevice(-1), lanesPerGpu(2) {}

ConvertToZDW::~ConvertToZDW() {}

const char* ConvertToZDW::compressorExtension() const {
  switch (compressor) {
    case GZIP: return ".gz";
    case BZIP2: return ".bz2";
    case XZ: case FXZ: return ".xz";
    case ZSTD: return ".zst";
  }
  return "";
}

const char* ConvertToZDW::compressorCommand() const {
  switch (compressor) {
    case GZIP: return "gzip";
    case BZIP2: return "bzip2";
    case XZ: return "xz";
    case FXZ: return "fxz";
    case ZSTD: return "zstd";
  }
  return "";
}

// .desc.sql -> column names, type ids, char sizes.  The matching is prefix based and order matters
// (ConvertToZDW.cpp:91-162): varchar(N); char(1) -> CHAR, char(2) -> CHAR_2, any other char(N) ->
// VARCHAR; text; tinytext; mediumtext; longtext; datetime; decimal (also one character in); everything else is an
// integer, signed unless the line mentions "unsigned": tinyint / smallint / bigint, and LONG for the rest (int,
// mediumint, float, double, timestamp, ...).  Lines starting with "Field" (any case) are skipped; a line without
// a tab is an error.  Lines are cut at 1023 bytes like the reference's fgets buffer.
bool ConvertToZDW::readDescFile(FILE* f, DescSchema& out) {
  char line[1024];
  while (fgets(line, sizeof(line), f)) {
    if (!strncasecmp(line, "Field", 5)) continue;
    char* tab = strchr(line, '\t');
    if (!tab) return false;
    *tab = 0;
    const char* type = tab + 1;
    out.names.push_back(line);
    int charSize = 0;
    unsigned char id;
    if (startsWith(type, "varchar")) {
      id = ZT_VARCHAR;
      charSize = atoi(type + 8);
    } else if (startsWith(type, "char")) {
      charSize = atoi(type + 5);
      id = charSize == 1 ? ZT_CHAR : charSize == 2 ? ZT_CHAR_2 : ZT_VARCHAR;
    } else if (startsWith(type, "text")) {
      id = ZT_TEXT;
    } else if (startsWith(type, "tinytext")) {
      id = ZT_TINYTEXT;
    } else if (startsWith(type, "mediumtext")) {
      id = ZT_MEDIUMTEXT;
    } else if (startsWith(type, "longtext")) {
      id = ZT_LONGTEXT;
    } else if (startsWith(type, "datetime")) {
      id = ZT_DATETIME;
    } else if (startsWith(type, "decimal") || (type[0] && startsWith(type + 1, "decimal"))) {
      id = ZT_DECIMAL;
    } else {
      const bool isSigned = strstr(type, "unsigned") == NULL;
      if (startsWith(type, "tinyint")) id = isSigned ? ZT_TINY_SIGNED : ZT_TINY;
      else if (startsWith(type, "smallint")) id = isSigned ? ZT_SHORT_SIGNED : ZT_SHORT;
      else if (startsWith(type, "bigint")) id = isSigned ? ZT_LONGLONG_SIGNED : ZT_LONGLONG;
      else id = isSigned ? ZT_LONG_SIGNED : ZT_LONG;
    }
    out.types.push_back(id);
    out.charSizes.push_back(charSize);
  }
  return true;
}

// keys may not contain '=' or a newline, values no newline (ConvertToZDW.cpp:226-237)
bool ConvertToZDW::metadataIsValid(const map<string, string>& metadata) {
  for (map<string, string>::const_iterator it = metadata.begin(); it != metadata.end(); ++it) {
    if (it->first.find_first_of("=\n") != string::npos) return false;
    if (it->second.find('\n') != string::npos) return false;
  }
  return true;
}

int ConvertToZDW::loadMetadataFile(const char* filepath, map<string, string>& metadata) {
  std::ifstream in(filepath);
  if (!in) return -1;
  string line;
  int lineNo = 0;
  while (std::getline(in, line)) {
    ++lineNo;
    if (line.empty()) continue;
    const size_t eq = line.find('=');
    if (eq == string::npos) return lineNo;
    metadata[line.substr(0, eq)] = line.substr(eq + 1);
  }
  return 0;
}

// version u16 | metadata length u32 | key\0value\0... | name\0...\0 | type[nc] | charSize u16[nc]
// (ConvertToZDW.cpp:673-737)
void ConvertToZDW::writeFileHeader(FILE* out, const DescSchema& schema, const map<string, string>& metadata) {
  string h;
  const uint16_t version = (uint16_t)CONVERT_ZDW_CURRENT_VERSION;
  h.append(reinterpret_cast<const char*>(&version), 2);
  uint32_t metaLen = 0;
  for (map<string, string>::const_iterator it = metadata.begin(); it != metadata.end(); ++it)
    metaLen += (uint32_t)(it->first.size() + it->second.size() + 2);
  h.append(reinterpret_cast<const char*>(&metaLen), 4);
  for (map<string, string>::const_iterator it = metadata.begin(); it != metadata.end(); ++it) {
    h.append(it->first).push_back('\0');
    h.append(it->second).push_back('\0');
  }
  for (size_t c = 0; c < schema.names.size(); ++c) h.append(schema.names[c]).push_back('\0');
  h.push_back('\0');
  h.append(reinterpret_cast<const char*>(schema.types.data()), schema.types.size());
  for (size_t c = 0; c < schema.charSizes.size(); ++c) {
    const uint16_t cs = (uint16_t)schema.charSizes[c];
    h.append(reinterpret_cast<const char*>(&cs), 2);
  }
  fwrite(h.data(), 1, h.size(), out);
}

// Round-trips the new file through unconvertDWfile (the one next to this executable) and compares with the source
// bytes (ConvertToZDW.cpp:166-223).  The .desc.sql files are not compared.
ConvertToZDW::ERR_CODE ConvertToZDW::validate(const char* zdwFile, const vector<string>& srcFiles, const char* exeName,
                                              const char* outputDir) {
  if (!bQuiet) statusOutput(INFO, "Unconverting %s back for validation...\n", zdwFile);
  string dir = exeName;
  const size_t slash = dir.rfind('/');
  dir.resize(slash == string::npos ? 0 : slash + 1);
  string decode = dir + "unconvertDWfile -q - ";
  if (outputDir) decode += string("-d ") + outputDir + " ";
  decode += zdwFile;
  string cmd;
  if (bStreamingInput) {
    string zcat = "zcat ";
    for (size_t i = 0; i < srcFiles.size(); ++i) zcat += srcFiles[i] + " ";
    cmd = "/bin/bash -c \"cmp <(" + decode + ") <(" + zcat + ")\"";
  } else if (bTrimTrailingSpaces) {
    cmd = "/bin/bash -c \"cmp <(" + decode + ") <(" + dir + "trim_spaces " + srcFiles[0] + ")\"";
  } else {
    cmd = decode + " | cmp " + srcFiles[0];
  }
  if (!bQuiet) statusOutput(INFO, "VALIDATION COMMAND: %s\n", cmd.c_str());
  return system(cmd.c_str()) == 0 ? OK : FILES_DIFFER;
}

// ---- several encode workers (the reference's block loop is ConvertToZDW.cpp:765-894) ----------------
// The windows of the file are cut on the host first (block_pipeline.h), then dealt out: worker w takes the next window
// that nobody has, reads it straight from the file into its own pinned buffer (so the reads run side by side too),
// encodes it on its GPU and hands the block to the calling thread, which puts the blocks out in file order with isLast
// and the cumulative longestLine patched (:841-842, :965).
struct ConvertToZDW::ParallelOutcome {
  uint64_t totalRows;
  int blocks;
  bool wrongColumns;
  uint32_t badRow;
  ParallelOutcome() : totalRows(0), blocks(0), wrongColumns(false), badRow(0) {}
};

namespace {
struct EncodedBlock {
  int rc;
  std::string err;
  vector<unsigned char> bytes;
  uint32_t nrows, longest, badRow, idxSize;
  uint64_t dictBytes, dictEntries;
  bool skipped;
  EncodedBlock() : rc(ZDWB_OK), nrows(0), longest(0), badRow(0), idxSize(0), dictBytes(0), dictEntries(0), skipped(false) {}
};
uint32_t readU32(const unsigned char* p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
}
}  // namespace

ConvertToZDW::ERR_CODE ConvertToZDW::encodeWindowsParallel(FILE* in, size_t windowBytes, const DescSchema& schema,
                                                           const char* filestub, const char* exeName, AsyncWriter& writer,
                                                           ParallelOutcome& outcome) {
  const int fd = fileno(in);
  struct stat st;
  if (fstat(fd, &st) != 0) return MISSING_SQL_FILE;
  vector<FileWindow> wins;
  size_t cap = windowBytes;
  if (!planFileWindows(fd, (uint64_t)st.st_size, windowBytes, MAX_WINDOW_BYTES, wins, &cap)) {
    statusOutput(ERROR, "%s: a single row exceeds %zu bytes (or the input could not be read)\n", exeName, (size_t)MAX_WINDOW_BYTES);
    return UNKNOWN_ERROR;
  }
  const double tStart = nowSeconds();
  if (hostTiming()) fprintf(stderr, "[zdw host] %zu windows planned (window %zu bytes)\n", wins.size(), cap);

  vector<int> workers;  // CUDA device of every worker
  {
    vector<int> devices = gpuList;
    if (devices.empty()) devices.push_back(GpuSession::resolve(gpuDevice));
    const int lanes = std::max(1, lanesPerGpu);
    for (int l = 0; l < lanes; ++l)  // lane-major: the first windows go to different devices
      for (size_t d = 0; d < devices.size(); ++d) workers.push_back(devices[d]);
    if (workers.size() > wins.size()) workers.resize(std::max<size_t>(1, wins.size()));
  }
  zdwb_schema sch;
  sch.ncols = (uint32_t)schema.types.size();
  sch.types = schema.types.data();

  OrderedResults<EncodedBlock> results(wins.size());
  std::atomic<size_t> next(0);
  std::atomic<bool> cancel(false);
  const size_t ahead = workers.size() + 2;
  const bool trim = bTrimTrailingSpaces;
  auto work = [&](int device) {
    GpuSession session;
    std::string fatal;
    int fatalRc = ZDWB_OK;
    const double tw0 = nowSeconds();
    double tOpen = 0, tRead = 0, tEnc = 0;
    // The first window of a worker is read into plain memory while its CUDA context comes up (about a second, and plain
    // memory needs no CUDA); every later one goes from the file straight to the device, 8 MiB at a time through the
    // context's pinned ring (zdwb_fd_to_device) - no window-sized host buffer, pinned (0.45 s per GiB to allocate) or
    // pageable (2.5 GB/s through the driver's staging).
    session.prefetch(device);
    char* buf = NULL;
    bool opened = false, sessionReady = false;
    for (;;) {
      const size_t k = next.fetch_add(1);
      if (k >= wins.size()) break;
      results.waitTurn(k, ahead);
      EncodedBlock r;
      const FileWindow& w = wins[k];
      const bool first = !sessionReady;
      size_t got = 0;
      if (first && !cancel) {
        buf = static_cast<char*>(malloc(w.len + 64));
        if (!buf) {
          fatal = "window allocation failed";
          fatalRc = ZDWB_ERR_OOM;
        }
        const double tr = nowSeconds();
        while (buf && got < w.len) {
          const ssize_t n = pread(fd, buf + got, w.len - got, (off_t)(w.offset + got));
          if (n < 0 && errno == EINTR) continue;
          if (n <= 0) break;
          got += (size_t)n;
        }
        tRead += nowSeconds() - tr;
      }
      if (!sessionReady) {
        const double to = nowSeconds();
        opened = session.open(device);
        tOpen = nowSeconds() - to;
        sessionReady = true;
      }
      if (!opened && fatalRc == ZDWB_OK) {
        fatal = "no usable CUDA device (" + session.lastError() + "); this build has no CPU path";
        fatalRc = ZDWB_ERR_NO_DEVICE;
      }
      if (fatalRc != ZDWB_OK) {
        r.rc = fatalRc;
        r.err = fatal;
        cancel = true;
      } else if (cancel) {
        r.skipped = true;
      } else {
        zdwb_encode_opts eo;
        memset(&eo, 0, sizeof(eo));
        eo.trim_trailing_spaces = trim ? 1 : 0;
        eo.more_input_follows = w.more ? 1 : 0;
        zdwb_block_out blk;
        memset(&blk, 0, sizeof(blk));
        const double te = nowSeconds();
        if (first) {
          if (got != w.len) {
            r.rc = ZDWB_ERR_BAD_ARG;
            r.err = "short read of the input file";
          } else {
            r.rc = zdwb_encode_block(session.get(), &sch, buf, w.len, &eo, &blk);
          }
        } else {
          const void* dev = NULL;
          r.rc = zdwb_fd_to_device(session.get(), fd, (long long)w.offset, w.len, &dev);
          if (r.rc == ZDWB_OK) {
            eo.input_on_device = 1;
            r.rc = zdwb_encode_block(session.get(), &sch, dev, w.len, &eo, &blk);
          }
        }
        if (r.rc != ZDWB_OK && r.err.empty()) r.err = zdwb_last_error(session.get());
        tEnc += nowSeconds() - te;
        r.badRow = blk.bad_row;
        if (r.rc == ZDWB_OK && blk.nrows && w.more && blk.tsv_consumed != w.consumed) {
          r.rc = ZDWB_ERR_BAD_ARG;  // the host's cut and the GPU's disagree: never write such a file
          r.err = "internal error: window cut mismatch (host " + std::to_string(w.consumed) + ", GPU " +
                  std::to_string((unsigned long long)blk.tsv_consumed) + ")";
        }
        if (r.rc == ZDWB_OK) {
          r.bytes.assign(blk.bytes, blk.bytes + blk.len);
          r.nrows = blk.nrows;
          r.longest = blk.longest_line;
          r.idxSize = blk.dict_index_size;
          r.dictBytes = blk.dict_bytes;
          r.dictEntries = blk.dict_entries;
        } else {
          cancel = true;
        }
      }
      if (buf) {  // (only the first window used it)
        free(buf);
        buf = NULL;
      }
      results.put(k, std::move(r));
    }
    if (hostTiming())
      fprintf(stderr, "[zdw host] encode worker on device %d: open %.3f s, first read %.3f s, upload + encode %.3f s, total %.3f s\n",
              device, tOpen, tRead, tEnc, nowSeconds() - tw0);
  };
  vector<std::thread> threads;
  for (size_t w = 0; w < workers.size(); ++w) threads.push_back(std::thread(work, workers[w]));

  ERR_CODE res = OK;
  uint32_t longestLine = 0;
  vector<unsigned char> pending;  // the previous block, held back until we know whether another one follows
  for (size_t k = 0; k < wins.size(); ++k) {
    EncodedBlock r = results.take(k);
    if (res != OK || outcome.wrongColumns || r.skipped) continue;  // (drain: the workers must get their turns)
    if (r.rc == ZDWB_ERR_WRONG_COLUMNS) {
      outcome.wrongColumns = true;
      outcome.badRow = r.badRow;
      continue;
    }
    if (r.rc == ZDWB_ERR_OOM) {
      statusOutput(ERROR, "Not enough memory to run %s\n", exeName);
      res = OUT_OF_MEMORY;
      continue;
    }
    if (r.rc != ZDWB_OK) {
      statusOutput(ERROR, "%s: GPU encode failed: %s\n", exeName, r.err.c_str());
      res = UNKNOWN_ERROR;
      continue;
    }
    if (r.nrows == 0) continue;  // only blank lines / an unterminated tail
    ++outcome.blocks;
    if (!bQuiet) {
      if (outcome.blocks == 1) statusOutput(INFO, "\nProcessing %s\n", filestub);
      else statusOutput(INFO, "\nProcessing block %d of %s (%llu rows so far)\n", outcome.blocks, filestub, (unsigned long long)outcome.totalRows);
      statusOutput(INFO, "Compiling unique values\n");
      statusOutput(INFO, "\r%u rows\n", r.nrows);
      statusOutput(INFO, "\nWriting dictionary:\n%u bytes being stored for %u unique entries.  Generating %d-byte offsets...\n",
                   (unsigned)r.dictBytes, (unsigned)r.dictEntries, (int)r.idxSize);
      statusOutput(INFO, "\nWriting rows\n");
    }
    if (!pending.empty()) {
      pending[8] = 0;  // another block follows (:841-842)
      writer.push(pending);
    }
    // every block was encoded on its own (prev_longest_line = 0): longestLine is cumulative over the file (:965)
    longestLine = std::max(longestLine, readU32(r.bytes.data() + 4));
    memcpy(r.bytes.data() + 4, &longestLine, 4);
    pending.swap(r.bytes);
    outcome.totalRows += r.nrows;
    if (!bQuiet) statusOutput(INFO, "\r%u\nDone with block %d -- cleaning up...\n", r.nrows, outcome.blocks);
  }
  for (size_t w = 0; w < threads.size(); ++w) threads[w].join();
  if (hostTiming()) fprintf(stderr, "[zdw host] all blocks encoded after %.3f s\n", nowSeconds() - tStart);
  if (res == OK && !outcome.wrongColumns && !pending.empty()) {
    pending[8] = 1;
    writer.push(pending);
  }
  return res;
}

ConvertToZDW::ERR_CODE ConvertToZDW::processFile(FILE* in, const char* filestub, const DescSchema& schema,
                                                 const bool bValidate, const char* exeName, const char* outputDir,
                                                 const char* zArgs, const map<string, string>& metadata) {
  if (!metadataIsValid(metadata)) {
    statusOutput(ERROR, "Invalid metadata parameter\n");
    return BAD_METADATA_PARAM;
  }
  gpu.prefetch(gpuList.empty() ? gpuDevice : gpuList[0]);  // CUDA start-up (a few hundred milliseconds) runs beside the first read

  // <outputDir or source dir>/<base>.zdw<ext>, written as <base>.creating.zdw<ext> and renamed on success (:629-657)
  string basePath;
  if (!outputDir) {
    basePath = filestub;
  } else {
    const char* base = strrchr(filestub, '/');
    basePath = string(outputDir) + "/" + (base ? base + 1 : filestub);
  }
  const string finalName = basePath + ".zdw" + compressorExtension();
  const string tempName = basePath + ".creating.zdw" + compressorExtension();

  ERR_CODE res = OK;
  vector<string> srcFiles;  // what validation compares against
  FILE* tee = NULL;
  FILE* out = NULL;
  string teeName;
  if (bStreamingInput) {
    // streamed input is kept (gzipped) for validation, like the reference's per-block temp files (:786-799)
    if (bValidate) {
      teeName = basePath + ".tmp.0.gz";
      tee = popen(("gzip > " + teeName).c_str(), "w");
      if (!tee) return CANT_OPEN_TEMP_FILE;
      srcFiles.push_back(teeName);
    }
  } else {
    srcFiles.push_back(string(filestub) + "." + getInputFileExtension());
  }

  // The input window.  A regular file that fits one window is read into plain memory while the CUDA context comes up;
  // a longer one that is not dealt to encode workers (and a pipe) goes through two window buffers: while the GPU encodes
  // one window a helper thread reads the next.
  ReadAheadInput win;
  AsyncWriter writer;
  const bool noReadAhead = getenv("ZDW_NO_READAHEAD") != NULL;  // measurement aid (tools/cli_timing.py --stream): read, then encode
  size_t windowBytes = std::min(std::max(blockBytes, (size_t)1 << 16), MAX_WINDOW_BYTES);
  bool oneWindow = false;
  bool regularInput = false;  // (only a regular file can be cut into windows up front and dealt to encode workers)
  if (!bStreamingInput) {
    struct stat st;
    regularInput = fstat(fileno(in), &st) == 0 && S_ISREG(st.st_mode);
    if (regularInput) posix_fadvise(fileno(in), 0, 0, POSIX_FADV_SEQUENTIAL);  // read front to back, once per byte
    if (regularInput && (unsigned long long)st.st_size < windowBytes) {  // (0 = empty, or a file system without sizes)
      oneWindow = true;
      windowBytes = std::max((size_t)st.st_size + 1, (size_t)4096);  // + 1: the read that finds the end of the file
    }
  }
  // Several windows of a regular file, cut by their size only: whole blocks go to several encode workers (and GPUs).
  const size_t nWorkers = std::max<size_t>(1, gpuList.size()) * (size_t)std::max(1, lanesPerGpu);
  const bool parallel = regularInput && !oneWindow && rowsPerBlock == 0 && blockPlan.empty() && heapBlocks == 0 && nWorkers > 1;
  if (oneWindow) {
    const size_t configured = std::min(std::max(blockBytes, (size_t)1 << 16), MAX_WINDOW_BYTES);
    if (!win.open(in, NULL, windowBytes, plainAlloc, plainFree)) return OUT_OF_MEMORY;
    win.fill();
    if (!win.eof()) {  // longer than stat() said (it grew, or the file system does not report sizes): the usual window
      windowBytes = configured;
      if (!win.widen(windowBytes)) return OUT_OF_MEMORY;
    }
  }
  // Several windows, one after the other (a pipe, or one worker): plain memory as well - the library moves it to the
  // device through its pinned ring at nearly the speed of a pinned window, a window-sized pinned buffer takes 0.45 s per
  // GiB to allocate, and above all plain memory needs no CUDA: the first window is read (the producer of a pipe keeps
  // running) while the context comes up.
  if (!parallel && !oneWindow) {
    if (!win.open(in, NULL, windowBytes, plainAlloc, plainFree)) {
      if (tee) {
        pclose(tee);
        unlink(teeName.c_str());
      }
      return OUT_OF_MEMORY;
    }
    win.fill();
  }
  if (!parallel && !gpu.open(gpuDevice)) {
    statusOutput(ERROR, "%s: no usable CUDA device (%s); this build has no CPU path\n", exeName, gpu.lastError().c_str());
    if (tee) {
      pclose(tee);
      unlink(teeName.c_str());
    }
    return UNKNOWN_ERROR;
  }

  string cmd = compressorCommand();
  if (zArgs) cmd += string(" ") + zArgs;
  cmd += " > " + tempName;
  out = popen(cmd.c_str(), "w");
  if (!out) {
    statusOutput(ERROR, "Could not open the process '%s' for writing!\n", cmd.c_str());
    if (tee) {
      pclose(tee);
      unlink(teeName.c_str());
    }
    return FILE_CREATION_ERR;
  }
  writeFileHeader(out, schema, metadata);
  writer.start(out);  // from here on the pipe belongs to the writer thread until writer.finish()

  zdwb_schema sch;
  sch.ncols = (uint32_t)schema.types.size();
  sch.types = schema.types.data();

  uint64_t totalRows = 0;
  uint32_t longestLine = 0;  // 0 = the 16 KiB start value (:965)
  int blocks = 0;
  bool wrongColumns = false;
  vector<unsigned char> pending;  // the previous block, held back until we know whether another one follows
  if (parallel) {
    ParallelOutcome po;
    res = encodeWindowsParallel(in, windowBytes, schema, filestub, exeName, writer, po);
    totalRows = po.totalRows;
    blocks = po.blocks;
    if (po.wrongColumns) {
      statusOutput(ERROR, "\nRow %u had the problem\n", po.badRow);
      wrongColumns = true;
    }
    if (res != OK) goto Done;
    if (!wrongColumns && blocks == 0) statusOutput(ERROR, "Empty data file -- nothing to process\n");
  }
  try {
  for (; !parallel;) {
    win.fill();
    if (win.len() == 0 && win.eof()) break;
    ++blocks;
    if (!bQuiet) {
      if (blocks == 1) statusOutput(INFO, "\nProcessing %s\n", filestub);
      else statusOutput(INFO, "\nProcessing block %d of %s (%llu rows so far)\n", blocks, filestub, (unsigned long long)totalRows);
      statusOutput(INFO, "Compiling unique values\n");
    }
    zdwb_encode_opts eo;
    memset(&eo, 0, sizeof(eo));
    eo.trim_trailing_spaces = bTrimTrailingSpaces ? 1 : 0;
    eo.more_input_follows = win.eof() ? 0 : 1;
    eo.prev_longest_line = longestLine;
    eo.max_rows = rowsPerBlock;
    const bool planned = (size_t)(blocks - 1) < blockPlan.size();
    if (planned) {
      eo.max_rows = blockPlan[blocks - 1].first;
      eo.spill_cols = blockPlan[blocks - 1].second;
    } else if (eo.max_rows == 0) {
      eo.heap_blocks = heapBlocks;  // the reference's own cut (a window in which heap block K does not open is widened below)
    }
    // a block cut by the window (not by a row count) uses the window up to its last row break: read ahead - from a
    // pipe (-i) as well, so that its producer keeps running while the GPU encodes (the helper polls: an error path
    // never waits for a read that does not return)
    if (eo.max_rows == 0 && eo.heap_blocks == 0 && !noReadAhead) win.prefetch();
    zdwb_block_out blk;
    const int rc = zdwb_encode_block(gpu.get(), &sch, win.data(), win.len(), &eo, &blk);
    if (rc == ZDWB_OK && planned && !win.eof() && blk.rows_in_buffer <= eo.max_rows) {
      // the window must hold the planned rows and the complete row after them: widen it and try again
      if (windowBytes >= MAX_WINDOW_BYTES) {
        statusOutput(ERROR, "%s: block %d of the plan does not fit %zu bytes\n", exeName, blocks, (size_t)MAX_WINDOW_BYTES);
        res = UNKNOWN_ERROR;
        goto Done;
      }
      windowBytes = std::min(windowBytes * 2, MAX_WINDOW_BYTES);
      if (!win.widen(windowBytes)) {
        res = OUT_OF_MEMORY;
        goto Done;
      }
      --blocks;
      continue;
    }
    if (rc == ZDWB_ERR_WRONG_COLUMNS) {
      statusOutput(ERROR, "\nRow %u had the problem\n", blk.bad_row);  // one past the last good row (:810-812)
      wrongColumns = true;
      break;
    }
    if (rc == ZDWB_ERR_OOM) {
      statusOutput(ERROR, "Not enough memory to run %s\n", exeName);
      res = OUT_OF_MEMORY;
      goto Done;
    }
    if (rc != ZDWB_OK) {
      statusOutput(ERROR, "%s: GPU encode failed: %s\n", exeName, zdwb_last_error(gpu.get()));
      res = UNKNOWN_ERROR;
      goto Done;
    }
    if (blk.nrows == 0) {
      if (!win.eof()) {
        // not one complete row in the window: widen it and try again
        if (windowBytes >= MAX_WINDOW_BYTES) {
          statusOutput(ERROR, "%s: a single row exceeds %zu bytes\n", exeName, (size_t)MAX_WINDOW_BYTES);
          res = UNKNOWN_ERROR;
          goto Done;
        }
        windowBytes = std::min(windowBytes * 2, MAX_WINDOW_BYTES);
        if (!win.widen(windowBytes)) {
          res = OUT_OF_MEMORY;
          goto Done;
        }
        --blocks;
        continue;
      }
      --blocks;
      break;  // only blank lines / an unterminated tail were left
    }
    if (!bQuiet) {
      statusOutput(INFO, "\r%u rows\n", blk.nrows);
      statusOutput(INFO, "\nWriting dictionary:\n%u bytes being stored for %u unique entries.  Generating %d-byte offsets...\n",
                   (unsigned)blk.dict_bytes, (unsigned)blk.dict_entries, (int)blk.dict_index_size);
      statusOutput(INFO, "\nWriting rows\n");
    }
    if (!pending.empty()) {
      pending[8] = 0;  // another block follows (:841-842)
      writer.push(pending);  // the compressor works on it while the next block is read and encoded
    }
    pending.assign(blk.bytes, blk.bytes + blk.len);
    if (tee && !writeRowsForValidation(tee, win.data(), (size_t)blk.tsv_consumed, bTrimTrailingSpaces)) {
      res = CANT_OPEN_TEMP_FILE;
      goto Done;
    }
    longestLine = blk.longest_line;
    totalRows += blk.nrows;
    if (!bQuiet) statusOutput(INFO, "\r%u\nDone with block %d -- cleaning up...\n", blk.nrows, blocks);
    win.consume((size_t)blk.tsv_consumed);
  }
  } catch (const std::bad_alloc&) {  // (the window could not grow): clean up like any other failure
    res = OUT_OF_MEMORY;
  }
  if (res != OK) goto Done;
  if (wrongColumns) {
    // the reference returns straight out of processFile here: the pipe is left to the process exit and the
    // .creating file stays on disk (:810-812).  We close the pipe but keep the file.
    win.close();
    writer.finish();
    if (tee) pclose(tee);
    if (!teeName.empty()) unlink(teeName.c_str());
    pclose(out);
    return WRONG_NUM_OF_COLUMNS_ON_A_ROW;
  }
  if (!pending.empty()) {
    pending[8] = 1;
    writer.push(pending);
  } else if (!parallel) {
    statusOutput(ERROR, "Empty data file -- nothing to process\n");  // :824-835, result stays OK
  }
  win.close();  // no reader left on `in` / `tee`
  writer.finish();
  if (tee) {
    pclose(tee);
    tee = NULL;
  }
  {
    // a compressor that died or a full disk must not end in a rename of a truncated file
    const bool flushed = fflush(out) == 0;
    const int status = pclose(out);
    out = NULL;
    if (!writer.ok() || !flushed || status != 0) {
      statusOutput(ERROR, "%s: writing %s failed (compressor exit status %d)\n", exeName, tempName.c_str(), status);
      res = FILE_CREATION_ERR;
      goto Done;
    }
  }

  if (bValidate) {
    const ERR_CODE v = validate(tempName.c_str(), srcFiles, exeName, outputDir);
    if (v == OK) {
      if (!bQuiet) statusOutput(INFO, "%s GOOD\n", finalName.c_str());
    } else {
      statusOutput(INFO, "%s BAD\n", finalName.c_str());
      res = v;
    }
  }

Done:
  win.close();
  writer.finish();
  if (tee) pclose(tee);
  if (out) pclose(out);
  if (!teeName.empty()) unlink(teeName.c_str());
  if (res == OK) {
    if (!bQuiet) statusOutput(INFO, "Rows=%u\n", (unsigned)totalRows);  // grepped by callers (:928-930)
    if (rename(tempName.c_str(), finalName.c_str()) == 0) {
      if (!bQuiet) statusOutput(INFO, "Done\n");
    } else {
      res = FILE_CREATION_ERR;
      statusOutput(INFO, "Final create file failed -- you can use %s instead.\n", tempName.c_str());
    }
  } else {
    unlink(tempName.c_str());
  }
  return res;
}

ConvertToZDW::ERR_CODE ConvertToZDW::convertFile(const char* infile, const char* exeName, const bool bValidate, char* filestub,
                                                 const char* outputDir, const char* zArgs,
                                                 const map<string, string>& metadata) {
  // the stub is the input path cut at its first ".sql" (:970-976)
  strcpy(filestub, infile);
  char* dot = strstr(filestub, ".sql");
  if (!dot) return MISSING_SQL_FILE;
  *dot = 0;

  const string descName = string(filestub) + ".desc." + getInputFileExtension();
  FILE* desc = fopen(descName.c_str(), "r");
  if (!desc) return MISSING_DESC_FILE;
  DescSchema schema;
  const bool descOk = readDescFile(desc, schema);
  fclose(desc);
  if (!descOk) return DESC_FILE_MISSING_TYPE_INFO;

  // <stub>.metadata is only consulted when no pairs were passed in (:991-999)
  map<string, string> meta = metadata;
  if (meta.empty()) {
    const int r = loadMetadataFile((string(filestub) + ".metadata").c_str(), meta);
    if (r > 0) return BAD_METADATA_FILE;
  }

  FILE* in = stdin;
  if (!bStreamingInput) {
    in = fopen((string(filestub) + "." + getInputFileExtension()).c_str(), "r");
    if (!in) return MISSING_SQL_FILE;
  }
  ERR_CODE res = UNKNOWN_ERROR;
  try {
    res = processFile(in, filestub, schema, bValidate, exeName, outputDir, zArgs, meta);
  } catch (const std::bad_alloc&) {
    res = OUT_OF_MEMORY;
  }
  if (!bStreamingInput) fclose(in);
  return res;
}

}  // namespace zdw
}  // namespace adobe
