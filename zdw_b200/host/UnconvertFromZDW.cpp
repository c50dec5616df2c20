// UnconvertFromZDW.cpp -- see zdw/UnconvertFromZDW.h.  Reference behaviour cited as cplusplus/UnconvertFromZDW.cpp:<line>.
#include "zdw/UnconvertFromZDW.h"
#include "block_pipeline.h"

#include <assert.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <strings.h>
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <deque>
#include <sstream>

#include <unistd.h>

using std::map;
using std::set;
using std::string;
using std::vector;

namespace adobe {
namespace zdw {

const int UnconvertFromZDW_Base::UNCONVERT_ZDW_VERSION = 11;
const char UnconvertFromZDW_Base::UNCONVERT_ZDW_VERSION_TAIL[3] = "c";

const char UnconvertFromZDW_Base::ERR_CODE_TEXTS[ERR_CODE_COUNT + 1][30] = {
  "OK", "BAD_PARAMETER", "GZREAD_FAILED", "FILE_CREATION_ERR", "FILE_OPEN_ERR", "UNSUPPORTED_ZDW_VERSION_ERR",
  "ZDW_LONGER_THAN_EXPECTED_ERR", "UNEXPECTED_DESC_TYPE", "ROW_COUNT_ERR", "CORRUPTED_DATA_ERROR", "HEADER_NOT_READ_YET",
  "HEADER_ALREADY_READ_ERR", "AT_END_OF_FILE", "BAD_REQUESTED_COLUMN", "NO_COLUMNS_TO_OUTPUT", "PROCESSING_ERROR",
  "UNSUPPORTED_OPERATION", "METADATA_KEY_NOT_PRESENT", "Unknown error"};

namespace {

const char* const VIRTUAL_BASENAME = "virtual_export_basename";
const char* const VIRTUAL_ROW = "virtual_export_row";

struct CaseInsensitiveLess {
  bool operator()(const string& a, const string& b) const { return strcasecmp(a.c_str(), b.c_str()) < 0; }
};

string displayName(const string& file) { return file.empty() ? string("stdin") : file; }

bool endsWith(const string& s, const char* suffix) {
  const size_t n = strlen(suffix);
  return s.size() > n && s.compare(s.size() - n, n, suffix) == 0;
}

}  // namespace

namespace {
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
}  // namespace

ZDWException::ZDWException(const ERR_CODE errcode)
    : std::runtime_error(UnconvertFromZDW_Base::ERR_CODE_TEXTS[errcode]), code(errcode) {}

// ---------------------------------------------------------------------------------------------------------------
// input
// ---------------------------------------------------------------------------------------------------------------
namespace internal {

ZdwInput::ZdwInput()
    : fp(NULL), isPipe(false), buf(NULL), cap(0), len(0), pos(0), ended(false), eofSeen(false), consumedTotal(0) {}

ZdwInput::~ZdwInput() {
  if (fp && isPipe) pclose(fp);
  else if (fp && fp != stdin) fclose(fp);
  free(buf);
}

bool ZdwInput::openCommand(const string& cmd) {
  fp = popen(cmd.c_str(), "r");
  isPipe = true;
  return fp != NULL;
}

bool ZdwInput::openFile(const string& path) {
  fp = fopen(path.c_str(), "rb");
  isPipe = false;
  if (fp) setvbuf(fp, NULL, _IONBF, 0);  // reads go straight into the block buffer, in large pieces
  return fp != NULL;
}

void ZdwInput::openStdin() {
  fp = stdin;
  isPipe = false;
}

size_t ZdwInput::ensure(size_t n) {
  if (len - pos >= n || ended || !fp) return len - pos;
  if (pos && pos == len) pos = len = 0;
  if (pos && pos + n > cap) {  // the request does not fit behind the read position: compact first
    memmove(buf, buf + pos, len - pos);
    len -= pos;
    pos = 0;
  }
  // The buffer grows as the bytes arrive, not to `n` up front: `n` comes out of a block header, and a corrupt one must
  // not be able to ask for more memory than the file is long.
  while (len - pos < n && !ended) {
    if (len == cap) {
      const size_t want = cap ? cap * 2 : ((size_t)1 << 20);
      char* nb = static_cast<char*>(realloc(buf, want + 64));
      if (!nb) throw std::bad_alloc();
      buf = nb;
      cap = want;
    }
    const size_t got = fread(buf + len, 1, cap - len, fp);
    len += got;
    if (got == 0) ended = true;
  }
  return len - pos;
}

void ZdwInput::consume(size_t n) {
  n = std::min(n, len - pos);
  pos += n;
  consumedTotal += n;
}

void ZdwInput::noteEofProbe() {
  if (ensure(1) == 0) eofSeen = true;
}

// The reference ends with a one-byte dummy read and then asks eof() (UnconvertFromZDW.cpp:1823-1834, :1937-1944):
// a single stray byte after the last block is swallowed, two or more are "longer than expected".
void ZdwInput::finalDummyRead() {
  if (ensure(1) >= 1) consume(1);
  if (ensure(1) == 0) eofSeen = true;
}

}  // namespace internal

// ---------------------------------------------------------------------------------------------------------------
// output
// ---------------------------------------------------------------------------------------------------------------
BufferedOutput::~BufferedOutput() {
  {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [this]() { return !busy; });
    stop = true;
  }
  cv.notify_all();
  if (started) worker.join();
}

void BufferedOutput::writeLater(const void* data, size_t size) {
  std::unique_lock<std::mutex> lk(m);
  cv.wait(lk, [this]() { return !busy; });
  if (!size) return;
  if (!started) {
    started = true;
    worker = std::thread([this]() { run(); });
  }
  jobData = data;
  jobSize = size;
  busy = true;
  cv.notify_all();
}

bool BufferedOutput::waitIdle() {
  std::unique_lock<std::mutex> lk(m);
  cv.wait(lk, [this]() { return !busy; });
  return !failed;
}

void BufferedOutput::run() {
  std::unique_lock<std::mutex> lk(m);
  for (;;) {
    cv.wait(lk, [this]() { return busy || stop; });
    if (!busy) return;  // stop is only raised while idle
    const void* data = jobData;
    const size_t size = jobSize;
    lk.unlock();
    const bool ok = fwrite(data, 1, size, fp) == size;
    lk.lock();
    if (!ok) failed = true;
    busy = false;
    cv.notify_all();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// base
// ---------------------------------------------------------------------------------------------------------------
UnconvertFromZDW_Base::UnconvertFromZDW_Base(const string& fileName, const bool showStatus, const bool quiet, const bool testOnly,
                                             const bool descOnly)
    : exportFileLineLength(0), virtualLineLength(0), version(UNCONVERT_ZDW_VERSION), numLines(0), numColumnsInExportFile(0),
      numColumns(0), lastBlock(1), inFileName(fileName), inFileBaseName(GetBaseNameForInFile(fileName)), input(NULL),
      bOutputDescFileOnly(descOnly), bShowStatus(showStatus && !quiet), bQuiet(quiet), bTestOnly(testOnly),
      bOutputNonEmptyColumnHeader(false), bShowBasicStatisticsOnly(false), bFailOnInvalidColumns(true),
      bExcludeSpecifiedColumns(false), bOutputEmptyMissingColumns(false), indexForVirtualBaseNameColumn(IGNORE_COLUMN),
      indexForVirtualRowColumn(IGNORE_COLUMN), rowsRead(0), rowsBeforeBlock(0), statusOutput(NULL), eState(ZDW_BEGIN),
      gpuDevice(-1), lanesPerGpu(2), lastBlockBytes(0), blocksToSink(0) {
  if (inFileName.empty()) {
    input = new internal::ZdwInput();
    input->openStdin();
    return;
  }
  struct stat st;
  if (stat(inFileName.c_str(), &st) < 0) return;  // isReadOpen() stays false
  // the decompressor is chosen by suffix and stays an external process (reference :233-263)
  string cmd;
  if (endsWith(inFileName, ".gz")) cmd = "zcat " + inFileName + " 2>/dev/null";
  else if (endsWith(inFileName, ".bz2")) cmd = "bzip2 -d --stdout " + inFileName + " 2>/dev/null";
  else if (endsWith(inFileName, ".xz")) cmd = "xzcat " + inFileName;
  else if (endsWith(inFileName, ".zst")) cmd = "zstd -d --stdout " + inFileName + " 2>/dev/null";
  input = new internal::ZdwInput();
  if (cmd.empty()) input->openFile(inFileName);  // reference: popen("cat <file>") - same bytes, one copy and one process less
  else input->openCommand(cmd);
}

UnconvertFromZDW_Base::~UnconvertFromZDW_Base() { delete input; }

string UnconvertFromZDW_Base::getVersion() {
  std::ostringstream s;
  s << UNCONVERT_ZDW_VERSION << UNCONVERT_ZDW_VERSION_TAIL;
  return s.str();
}

// callers grep this line to detect a failed unconvert (reference :172-176)
void UnconvertFromZDW_Base::printError(const string& exe, const string& fileName) {
  statusOutput(ERROR, "%s: %s failed\n\n", !exe.empty() ? exe.c_str() : "UnconvertFromZDW", fileName.c_str());
}

size_t UnconvertFromZDW_Base::readBytes(void* dst, const size_t n, const bool haltOnError) {
  const size_t have = std::min(input->ensure(n), n);
  if (have) memcpy(dst, input->data(), have);
  input->consume(have);
  if (have != n) {
    if (have == 0) input->noteEofProbe();
    if (haltOnError) {
      printError(exeName, displayName(inFileName));
      throw ZDWException(GZREAD_FAILED);  // the one exception that crosses the API (reference :286-302)
    }
  }
  return have;
}

// <dir>/<name>.zdw[.ext] -> dir, name: the last ".zdw" and everything after it is cut (reference :98-133)
void UnconvertFromZDW_Base::splitDirAndBase(const string& file, string& dir, string& base) {
  const size_t slash = file.rfind('/');
  if (slash == string::npos) {
    dir = ".";
    base = file;
  } else {
    dir = file.substr(0, slash);
    base = file.substr(slash + 1);
  }
  const size_t z = base.rfind(".zdw");
  if (z != string::npos) base.resize(z);
}

string UnconvertFromZDW_Base::GetBaseNameForInFile(const string& file) {
  if (file.empty()) return string();
  string dir, base;
  splitDirAndBase(file, dir, base);
  return base;
}

size_t UnconvertFromZDW_Base::numOutputColumns() const {
  if (namesOfColumnsToOutput.empty()) return numColumns;
  size_t n = blankColumnNames.size();
  for (size_t c = 0; c < outputColumns.size(); ++c)
    if (outputColumns[c] != IGNORE_COLUMN) ++n;
  return n;
}

// ---- column selection (-c / -ci / -ce / -cx), reference :404-507 -------------------------------------------------
bool UnconvertFromZDW_Base::setNamesOfColumnsToOutput(const vector<string>& requested, COLUMN_INCLUSION_RULE rule) {
  namesOfColumnsToOutput.clear();
  bFailOnInvalidColumns = rule == FAIL_ON_INVALID_COLUMN || rule > PROVIDE_EMPTY_MISSING_COLUMNS;
  bExcludeSpecifiedColumns = rule == EXCLUDE_SPECIFIED_COLUMNS;
  bOutputEmptyMissingColumns = rule == PROVIDE_EMPTY_MISSING_COLUMNS;
  unsigned next = 0;
  for (size_t k = 0; k < requested.size(); ++k) {
    const string& name = requested[k];
    const bool added = namesOfColumnsToOutput.insert(std::make_pair(name, next)).second;
    if (!bExcludeSpecifiedColumns) {
      if (name == VIRTUAL_BASENAME) indexForVirtualBaseNameColumn = USE_VIRTUAL_COLUMN;
      if (name == VIRTUAL_ROW) indexForVirtualRowColumn = USE_VIRTUAL_COLUMN;
    }
    if (added) {
      ++next;
    } else {
      if (bFailOnInvalidColumns) return false;  // duplicate
      if (bOutputEmptyMissingColumns) blankColumnNames[next++] = name;
    }
  }
  return true;
}

bool UnconvertFromZDW_Base::setNamesOfColumnsToOutput(const string& csv, COLUMN_INCLUSION_RULE rule) {
  namesOfColumnsToOutput.clear();
  bFailOnInvalidColumns = rule == FAIL_ON_INVALID_COLUMN || rule > PROVIDE_EMPTY_MISSING_COLUMNS;
  bExcludeSpecifiedColumns = rule == EXCLUDE_SPECIFIED_COLUMNS;
  bOutputEmptyMissingColumns = rule == PROVIDE_EMPTY_MISSING_COLUMNS;
  // names are separated by commas and/or spaces; the case-insensitive duplicate rule applies to this overload only
  set<string, CaseInsensitiveLess> seen;
  unsigned next = 0;
  size_t at = csv.find_first_not_of(", ");
  while (at != string::npos) {
    size_t end = csv.find_first_of(", ", at);
    const string name = csv.substr(at, end == string::npos ? string::npos : end - at);
    if (seen.insert(name).second) {
      namesOfColumnsToOutput.insert(std::make_pair(name, next));
      ++next;
      if (!bExcludeSpecifiedColumns) {
        if (name == VIRTUAL_BASENAME) indexForVirtualBaseNameColumn = USE_VIRTUAL_COLUMN;
        else if (name == VIRTUAL_ROW) indexForVirtualRowColumn = USE_VIRTUAL_COLUMN;
      }
    } else {
      if (bFailOnInvalidColumns) return false;
      if (bOutputEmptyMissingColumns) blankColumnNames[next++] = name;
    }
    at = end == string::npos ? string::npos : csv.find_first_not_of(", ", end);
  }
  return true;
}

// ---- file header, reference :1030-1219 ----------------------------------------------------------------------------
ERR_CODE UnconvertFromZDW_Base::readHeader() {
  columnType.clear();
  columnCharSize.clear();
  if (!isReadOpen()) return FILE_OPEN_ERR;
  if (eState != ZDW_BEGIN) return HEADER_ALREADY_READ_ERR;
  if (!statusOutput) statusOutput = defaultStatusOutputCallback;
  if (!bOutputDescFileOnly) gpu.prefetch(gpuDevice);  // CUDA start-up overlaps the reading of the header and first block

  readBytes(&version, 2);
  if (version > UNCONVERT_ZDW_VERSION) return UNSUPPORTED_ZDW_VERSION_ERR;
  if (bShowBasicStatisticsOnly) statusOutput(INFO, "File version %d\n", (int)version);
  if (version < 9) {
    // v1-v8 use the tree dictionary / visitor tables the encoder stopped writing; not carried over
    statusOutput(ERROR, "%s: ZDW version %d files are not supported by this build (versions 9-11 are)\n",
                 exeName.empty() ? "UnconvertFromZDW" : exeName.c_str(), (int)version);
    return UNSUPPORTED_ZDW_VERSION_ERR;
  }

  metadata.clear();
  if (version >= 11) {
    ULONG metaSize = 0;
    readBytes(&metaSize, 4);
    if (bShowBasicStatisticsOnly) statusOutput(INFO, "Metadata block size = %u\n", metaSize);
    vector<char> block(metaSize + 1, 0);
    if (metaSize) readBytes(block.data(), metaSize);
    size_t at = 0;
    while (at < metaSize) {
      const string key(block.data() + at);
      at += key.size() + 1;
      const string value(at < metaSize ? block.data() + at : "");
      at += value.size() + 1;
      metadata[key] = value;
    }
  }

  // column names: name\0 ... \0
  columnNames.clear();
  for (;;) {
    string name;
    char ch;
    readBytes(&ch, 1);
    if (!ch) break;
    while (ch) {
      name.push_back(ch);
      readBytes(&ch, 1);
    }
    columnNames.push_back(name);
  }
  numColumnsInExportFile = (ULONG)columnNames.size();

  // virtual columns sit behind the file's columns (:1101-1111)
  if (UseVirtualExportBaseNameColumn()) {
    indexForVirtualBaseNameColumn = (int)columnNames.size();
    columnNames.push_back(VIRTUAL_BASENAME);
    virtualLineLength += (ULONG)inFileBaseName.size() + 1;
  }
  if (UseVirtualExportRowColumn()) {
    indexForVirtualRowColumn = (int)columnNames.size();
    columnNames.push_back(VIRTUAL_ROW);
    virtualLineLength += 20 + 1;  // digits of SIZE_MAX
  }
  numColumns = (ULONG)columnNames.size();

  // which column goes where (:1113-1190)
  outputColumns.assign(numColumns, namesOfColumnsToOutput.empty() ? 0 : IGNORE_COLUMN);
  map<string, unsigned, CaseInsensitiveLess> wanted(namesOfColumnsToOutput.begin(), namesOfColumnsToOutput.end());
  map<unsigned, unsigned> placed;  // output position -> file column
  unsigned nextOut = 0;
  for (unsigned c = 0; c < numColumns; ++c) {
    map<string, unsigned, CaseInsensitiveLess>::iterator hit = wanted.find(columnNames[c]);
    if (bExcludeSpecifiedColumns) {
      if (hit == wanted.end()) outputColumns[c] = (int)nextOut++;
    } else if (hit != wanted.end()) {
      outputColumns[c] = (int)hit->second;
      placed[hit->second] = c;
      wanted.erase(hit);
    }
  }
  if (!wanted.empty() && !bExcludeSpecifiedColumns) {
    if (bFailOnInvalidColumns) return BAD_REQUESTED_COLUMN;
    if (bOutputEmptyMissingColumns) {
      for (map<string, unsigned, CaseInsensitiveLess>::const_iterator it = wanted.begin(); it != wanted.end(); ++it)
        blankColumnNames[(int)it->second] = it->first;
    } else {
      if (placed.empty()) return NO_COLUMNS_TO_OUTPUT;
      unsigned k = 0;  // close the gaps the missing names left: [2,1,3,5] -> [1,0,2,3]
      for (map<unsigned, unsigned>::const_iterator it = placed.begin(); it != placed.end(); ++it, ++k)
        if (it->first != k) outputColumns[it->second] = (int)k;
    }
  }

  columnType.assign(numColumns, 0);
  readBytes(columnType.data(), numColumnsInExportFile);
  columnCharSize.assign(numColumns, 0);
  readBytes(columnCharSize.data(), (size_t)numColumnsInExportFile * 2);  // version >= 7 always holds here
  if (UseVirtualExportBaseNameColumn()) {
    columnType[indexForVirtualBaseNameColumn] = ZT_VIRTUAL_EXPORT_FILE_BASENAME;
    columnCharSize[indexForVirtualBaseNameColumn] = (USHORT)(inFileBaseName.size() + 1);
  }
  if (UseVirtualExportRowColumn()) {
    columnType[indexForVirtualRowColumn] = ZT_VIRTUAL_EXPORT_ROW;
    columnCharSize[indexForVirtualRowColumn] = 0;
  }
  setState(ZDW_PARSE_BLOCK_HEADER);
  return OK;
}

// ---- .desc.sql / schema / .metadata emitters, reference :510-732 ----------------------------------------------------
string UnconvertFromZDW_Base::getColumnDesc(const string& name, UCHAR type, size_t index, const string& sep,
                                            const string& delimiter) const {
  string text = name + sep;
  switch (type) {
    case ZT_VIRTUAL_EXPORT_FILE_BASENAME:
    case ZT_VARCHAR: {
      const int n = (index < columnCharSize.size() && columnCharSize[index]) ? columnCharSize[index] : 255;
      char tmp[32];
      snprintf(tmp, sizeof(tmp), "varchar(%d)", n);
      text += tmp;
      break;
    }
    case ZT_TEXT: text += "text"; break;
    case ZT_TINYTEXT: text += "tinytext"; break;
    case ZT_MEDIUMTEXT: text += "mediumtext"; break;
    case ZT_LONGTEXT: text += "longtext"; break;
    case ZT_DATETIME: text += "datetime"; break;
    case ZT_CHAR_2: text += "char(2)"; break;
    case ZT_VISID_LOW: case ZT_VISID_HIGH: case ZT_LONGLONG: text += "bigint(20) unsigned"; break;
    case ZT_CHAR: text += "char(1)"; break;
    case ZT_TINY: text += "tinyint(3) unsigned"; break;
    case ZT_SHORT: text += "smallint(5) unsigned"; break;
    case ZT_VIRTUAL_EXPORT_ROW: case ZT_LONG: text += "int(11) unsigned"; break;
    case ZT_TINY_SIGNED: text += "tinyint(3)"; break;
    case ZT_SHORT_SIGNED: text += "smallint(5)"; break;
    case ZT_LONG_SIGNED: text += "int(11)"; break;
    case ZT_LONGLONG_SIGNED: text += "bigint(20)"; break;
    case ZT_DECIMAL: text += "decimal(24,12)"; break;
    default: return string();
  }
  return text + delimiter;
}

vector<string> UnconvertFromZDW_Base::getDesc(const vector<string>& names, const string& sep, const string& delimiter) const {
  vector<string> lines(names.size() + blankColumnNames.size());
  for (size_t c = 0; c < names.size(); ++c) {
    const int at = namesOfColumnsToOutput.empty() ? (int)c : outputColumns[c];
    if (at == IGNORE_COLUMN) continue;
    lines[at] = getColumnDesc(names[c], columnType[c], c, sep, delimiter);
    if (lines[at].empty()) return vector<string>();
  }
  // requested-but-absent columns are described as plain text
  for (map<int, string>::const_iterator it = blankColumnNames.begin(); it != blankColumnNames.end(); ++it)
    lines[it->first] = getColumnDesc(it->second, ZT_TEXT, (size_t)-1, sep, delimiter);
  return lines;
}

ERR_CODE UnconvertFromZDW_Base::outputDesc(const vector<string>& names, FILE* to) {
  const vector<string> lines = getDesc(names, "\t", "\n");
  if (lines.empty() && !names.empty()) return UNEXPECTED_DESC_TYPE;
  for (size_t k = 0; k < lines.size(); ++k) fputs(lines[k].c_str(), to);
  return OK;
}

ERR_CODE UnconvertFromZDW_Base::outputDescToFile(const vector<string>& names, const string& outputDir, const char* filestub,
                                                 const char* ext) {
  const string path = outputDir + "/" + filestub + ".desc" + (ext ? ext : "");
  FILE* f = fopen(path.c_str(), "w");
  if (!f) {
    statusOutput(ERROR, "%s: Could not open %s for writing\n", exeName.c_str(), path.c_str());
    return FILE_CREATION_ERR;
  }
  const ERR_CODE rc = outputDesc(names, f);
  fclose(f);
  return rc;
}

ERR_CODE UnconvertFromZDW_Base::outputDescToStdOut(const vector<string>& names) { return outputDesc(names, stdout); }

ERR_CODE UnconvertFromZDW_Base::GetSchema(std::ostream& stream) {
  const vector<string> lines = getDesc(columnNames, " ", "");
  if (lines.empty() && !columnNames.empty()) return UNEXPECTED_DESC_TYPE;
  for (size_t k = 0; k < lines.size(); ++k) stream << (k ? ",\n" : "") << lines[k];
  return OK;
}

ERR_CODE UnconvertFromZDW_Base::outputMetadata(FILE* to) const {
  const set<string>& keys = metadataOptions.keys;
  if (!metadataOptions.bAllowMissingKeys)
    for (set<string>::const_iterator k = keys.begin(); k != keys.end(); ++k)
      if (metadata.find(*k) == metadata.end()) return METADATA_KEY_NOT_PRESENT;
  if (keys.empty()) {
    for (map<string, string>::const_iterator m = metadata.begin(); m != metadata.end(); ++m) {
      if (metadataOptions.bOnlyMetadataKeys) fprintf(to, "%s\n", m->first.c_str());
      else fprintf(to, "%s=%s\n", m->first.c_str(), m->second.c_str());
    }
  } else {
    for (set<string>::const_iterator k = keys.begin(); k != keys.end(); ++k) {
      if (metadataOptions.bOnlyMetadataKeys) {
        fprintf(to, "%s\n", k->c_str());
      } else {
        map<string, string>::const_iterator m = metadata.find(*k);
        fprintf(to, "%s=%s\n", k->c_str(), m != metadata.end() ? m->second.c_str() : "");
      }
    }
  }
  return OK;
}

ERR_CODE UnconvertFromZDW_Base::outputMetadataToFile(const string& outputDir, const char* filestub) const {
  const string path = outputDir + "/" + filestub + ".metadata";
  FILE* f = fopen(path.c_str(), "w");
  if (!f) {
    statusOutput(ERROR, "%s: Could not open %s for writing\n", exeName.c_str(), path.c_str());
    return FILE_CREATION_ERR;
  }
  const ERR_CODE rc = outputMetadata(f);
  fclose(f);
  return rc;
}

ERR_CODE UnconvertFromZDW_Base::outputMetadataToStdOut() const { return outputMetadata(stdout); }

// ---- blocks ---------------------------------------------------------------------------------------------------------
// "***ZDW BLOCK HEADER*** NON-EMPTY COLUMNS: a,b,..." (reference :1002-1024)
string UnconvertFromZDW_Base::getBlockHeaderString(const BlockInfo& info) const {
  string header = "***ZDW BLOCK HEADER*** NON-EMPTY COLUMNS: ";
  bool first = true;
  for (size_t c = 0; c < numColumns; ++c) {
    if (outputColumns[c] == IGNORE_COLUMN) continue;
    if (c < info.columnSize.size() && info.columnSize[c]) {
      if (!first) header += ",";
      first = false;
      header += columnNames[c];
    }
  }
  return header + "\n";
}

ERR_CODE UnconvertFromZDW_Base::peekBlock(BlockInfo& info) {
  const size_t nc = numColumnsInExportFile;
  if (input->ensure(10) < 10) {
    if (input->available() == 0) input->noteEofProbe();
    printError(exeName, displayName(inFileName));
    throw ZDWException(GZREAD_FAILED);
  }
  const unsigned char* p = reinterpret_cast<const unsigned char*>(input->data());
  memcpy(&info.numLines, p, 4);
  memcpy(&info.lineLength, p + 4, 4);
  info.last = p[8];
  const unsigned idxSize = p[9];
  if (idxSize > 4) return CORRUPTED_DATA_ERROR;
  numLines = info.numLines;
  exportFileLineLength = info.lineLength;
  lastBlock = info.last;
  if (bShowBasicStatisticsOnly) statusOutput(INFO, "Max line length = %lu\n", (unsigned long)exportFileLineLength);

  info.dictionarySize = 0;
  size_t statsAt = 10;
  if (idxSize) {
    if (input->ensure(10 + idxSize) < 10 + idxSize) {
      printError(exeName, displayName(inFileName));
      throw ZDWException(GZREAD_FAILED);
    }
    p = reinterpret_cast<const unsigned char*>(input->data());
    ULONG d = 0;
    memcpy(&d, p + 10, idxSize);
    info.dictionarySize = d;
    statsAt = 10 + idxSize + (size_t)d;
  }
  if (!bQuiet) statusOutput(INFO, "Reading %llu byte dictionary\n", (unsigned long long)info.dictionarySize);
  if (input->ensure(statsAt + nc) < statsAt + nc) {
    printError(exeName, displayName(inFileName));
    throw ZDWException(GZREAD_FAILED);
  }
  p = reinterpret_cast<const unsigned char*>(input->data());
  info.columnSize.assign(p + statsAt, p + statsAt + nc);
  info.columnSize.resize(numColumns, 0);  // virtual columns have no storage (:989-1000)
  size_t used = 0, valueBytes = 0;
  for (size_t c = 0; c < nc; ++c)
    if (info.columnSize[c]) {
      ++used;
      valueBytes += info.columnSize[c];
    }
  info.numSetColumns = (used + 7) / 8;
  info.maxRowBytes = info.numSetColumns + valueBytes;
  info.rowsOffset = statsAt + nc + 8 * used;
  // Buffer what the block will probably need - its rows cannot take more than numLines * maxRowBytes, but delta-coded
  // rows are usually several times smaller, so that bound would pull the following blocks (often the rest of the file)
  // into memory.  The first guess is the size of the previous block plus a quarter (blocks of a file look alike), or a
  // quarter of the bound; decodeBlock asks for more when the GPU reports the block as cut short.
  const double tRead0 = nowSeconds();
  const size_t upper = info.rowsOffset + (size_t)info.numLines * info.maxRowBytes + 1;
  size_t guess = lastBlockBytes ? lastBlockBytes + lastBlockBytes / 4 + 4096
                                : info.rowsOffset + (size_t)info.numLines * info.maxRowBytes / 4 + 65536;
  input->ensure(std::min(upper, std::max(guess, info.rowsOffset + 1)));
  if (hostTiming()) fprintf(stderr, "[zdw host] input buffered %.3f s (%zu bytes)\n", nowSeconds() - tRead0, input->available());
  return OK;
}

int UnconvertFromZDW_Base::decodeBytes(GpuSession& g, const void* data, size_t avail, bool atEnd, unsigned long long firstRow,
                                       unsigned char separator, bool wantRowOffsets, bool validateOnly, bool wantFlagCounts,
                                       bool skimOnly, zdwb_rows_out* out, bool outputOnDevice) const {
  zdwb_schema sch;
  sch.ncols = numColumnsInExportFile;
  sch.types = columnType.data();
  zdwb_decode_opts o;
  memset(&o, 0, sizeof(o));
  o.want_row_offsets = wantRowOffsets ? 1 : 0;
  o.at_end_of_file = atEnd ? 1 : 0;
  o.separator = separator;
  o.rownum_pos = -1;
  o.validate_only = validateOnly ? 1 : 0;
  o.want_flag_counts = wantFlagCounts ? 1 : 0;
  o.skim_only = skimOnly ? 1 : 0;
  o.output_on_device = outputOnDevice ? 1 : 0;
  o.first_row_number = firstRow;
  vector<int32_t> map32;
  zdwb_fill fill;
  if (!namesOfColumnsToOutput.empty()) {
    map32.assign(outputColumns.begin(), outputColumns.begin() + numColumnsInExportFile);
    if (map32.empty()) map32.push_back(-1);
    o.out_col = map32.data();
    o.n_out = (uint32_t)numOutputColumns();
    if (UseVirtualExportBaseNameColumn() && outputColumns[indexForVirtualBaseNameColumn] != IGNORE_COLUMN) {
      fill.pos = (uint32_t)outputColumns[indexForVirtualBaseNameColumn];
      fill.len = (uint32_t)inFileBaseName.size();
      fill.text = inFileBaseName.c_str();
      o.fills = &fill;
      o.n_fills = 1;
    }
    if (UseVirtualExportRowColumn() && outputColumns[indexForVirtualRowColumn] != IGNORE_COLUMN)
      o.rownum_pos = outputColumns[indexForVirtualRowColumn];
  }
  return zdwb_decode_block(g.get(), &sch, data, avail, &o, out);
}

ERR_CODE UnconvertFromZDW_Base::decodeBlock(const BlockInfo& info, unsigned char separator, bool wantRowOffsets, bool validateOnly,
                                            bool wantFlagCounts, zdwb_rows_out* out, GpuSession* session, bool skimOnly) {
  GpuSession& g = session ? *session : gpu;
  const double tOpen0 = nowSeconds();
  const bool opened = g.open(gpuList.empty() ? gpuDevice : gpuList[0]);
  if (hostTiming()) fprintf(stderr, "[zdw host] gpu.open %.3f s\n", nowSeconds() - tOpen0);
  if (!opened) {
    statusOutput(ERROR, "%s: no usable CUDA device (%s); this build has no CPU path\n",
                 exeName.empty() ? "UnconvertFromZDW" : exeName.c_str(), g.lastError().c_str());
    return PROCESSING_ERROR;
  }
  int rc;
  for (;;) {
    const double tDec0 = nowSeconds();
    rc = decodeBytes(g, input->data(), input->available(), input->sourceEnded(), rowsBeforeBlock + 1, separator, wantRowOffsets,
                     validateOnly, wantFlagCounts, skimOnly, out);
    if (hostTiming())
      fprintf(stderr, "[zdw host] zdwb_decode_block%s %.3f s (%zu bytes available, %llu bytes out)\n", skimOnly ? " (skim)" : "",
              nowSeconds() - tDec0, input->available(), (unsigned long long)out->len);
    if (rc != ZDWB_ERR_TRUNCATED || input->sourceEnded()) break;
    // the block is longer than what was buffered for it: read on (at least as much again) and try once more
    const size_t have = input->available();
    if (input->ensure(have + std::max<size_t>(have, (size_t)1 << 20)) == have && !input->sourceEnded()) break;
  }
  (void)info;
  switch (rc) {
    case ZDWB_OK:
      if (out->consumed) lastBlockBytes = (size_t)out->consumed;
      return OK;
    case ZDWB_ERR_CORRUPT: return CORRUPTED_DATA_ERROR;  // reference :1364-1365
    case ZDWB_ERR_ROW_COUNT:
    case ZDWB_ERR_TRUNCATED:
      printError(exeName, displayName(inFileName));
      // (the reference prints the rows it had unpacked when the data ran out; here a block is decoded as a whole or not
      // at all, so the count of a block that is cut short is always 0 and none of its rows are written)
      statusOutput(INFO, "Rows unpacked (%u) does not match expected (%u)\n\n", 0u, numLines);  // :1597-1605
      return ROW_COUNT_ERR;
    default:
      statusOutput(ERROR, "%s: GPU decode failed: %s\n", exeName.empty() ? "UnconvertFromZDW" : exeName.c_str(),
                   zdwb_last_error(g.get()));
      return PROCESSING_ERROR;
  }
}

// One block to a file-like sink (reference parseNextBlock, :1472-1633).
template <typename T>
ERR_CODE UnconvertFromZDW<T>::parseNextBlock(T& sink) {
  typename UnconvertFromZDW_Base::BlockInfo info;
  ERR_CODE rc = this->peekBlock(info);
  if (rc != OK) return rc;
  this->rowsRead = 0;
  if (this->bOutputNonEmptyColumnHeader) {
    const string header = this->getBlockHeaderString(info);
    sink.write(header.data(), header.size());
  }
  if (!this->bQuiet) this->statusOutput(INFO, "Reading %u rows\n", this->numLines);

  const bool scanOnly = this->bTestOnly || (this->bShowBasicStatisticsOnly && !this->isLastBlock());
  zdwb_rows_out rows;
  memset(&rows, 0, sizeof(rows));
  if (scanOnly) {
    rc = this->decodeBlock(info, '\t', false, true, this->bShowBasicStatisticsOnly, &rows);
    if (rc != OK) return rc;
    this->rowsRead = this->numLines;
    this->input->consume((size_t)rows.consumed);
  } else if (!this->bShowBasicStatisticsOnly) {
    // Blocks alternate between two GPU contexts: the rows of this block stay valid in theirs while the sink's writer
    // thread puts them out and the next block is read (the decompressor pipe drains meanwhile) and decoded in the
    // other context.  The sink has one block in flight, so a context's rows are on disk before it is used again.
    GpuSession* session = (this->blocksToSink++ & 1) ? &this->gpu2 : &this->gpu;
    rc = this->decodeBlock(info, '\t', false, false, false, &rows, session);
    if (rc != OK) return rc;
    sink.writeLater(rows.tsv, rows.len);
    this->rowsRead = this->numLines;
    this->input->consume((size_t)rows.consumed);
    if (this->bShowStatus) this->statusOutput(INFO, "\r%u\n", this->rowsRead);
  }
  this->rowsBeforeBlock += this->numLines;

  if (this->bShowBasicStatisticsOnly && rows.flag_counts) {  // -s: equality-bit statistics (:1608-1624)
    unsigned long long total = 0;
    for (uint32_t u = 0; u < rows.ncols_used; ++u) total += rows.flag_counts[u];
    if (total) {
      this->statusOutput(INFO,
                         "Equality delta bits set: %llu (%0.1f%%) (rows=%u, columns=%u, bit vector width=%ld bytes, non-empty "
                         "columns=%zu (%0.1f%%)\n",
                         total, total * 100 / float((double)this->numLines * info.numSetColumns * 8), this->numLines,
                         this->numColumnsInExportFile, (long)info.numSetColumns, (size_t)rows.ncols_used,
                         rows.ncols_used * 100 / float(this->numColumnsInExportFile));
      for (uint32_t u = 0; u < rows.ncols_used; ++u) this->statusOutput(INFO, "%u ", (unsigned)rows.flag_counts[u]);
      this->statusOutput(INFO, "\n");
    }
  }
  if (this->isLastBlock() && !this->bQuiet && !this->bShowBasicStatisticsOnly)
    this->statusOutput(INFO, "%s %s\n\n", displayName(this->inFileName).c_str(), this->bTestOnly ? "tested good" : "uncompressed");
  return OK;
}

// ---- several decode workers (SURVEY 8(e) "Decode"; the reference's block loop is :1814-1820) -------------------------
// The calling thread walks the file: block header, then a SKIM of the block on its own context - the row-boundary
// kernels only - which yields the block's length, i.e. where the next block starts.  The block's bytes go to a worker
// (own context, any device), which decodes them and puts the rows out when it is the block's turn: in order through the
// FILE* for pipes, or side by side with pwrite() at the block's offset when the sink is a regular file (a block's offset
// is known as soon as the blocks in front of it have been decoded, not written).
namespace {

struct DecodeJob {
  size_t seq;
  vector<char> bytes;  // the whole block
  unsigned long long firstRow;
  ULONG numLines;
  string prefix;       // --non-empty-column-header line
};

}  // namespace

template <typename T>
ERR_CODE UnconvertFromZDW<T>::decodeBlocksFanOut(T& sink) {
  sink.waitIdle();
  OrderedSink ordered(sink.file());
  vector<int> workers;
  {
    vector<int> devices = this->gpuList;
    if (devices.empty()) devices.push_back(GpuSession::resolve(this->gpuDevice));
    for (int l = 0; l < std::max(1, this->lanesPerGpu); ++l)
      for (size_t d = 0; d < devices.size(); ++d) workers.push_back(devices[d]);
  }
  std::mutex qm;
  std::condition_variable qcv;
  std::deque<DecodeJob> queue;
  bool closed = false;
  std::atomic<int> firstError(OK);
  string workerErr;
  const size_t maxQueued = workers.size() + 1;

  auto work = [&](int device) {
    GpuSession session;
    const bool opened = session.open(device);
    for (;;) {
      DecodeJob job;
      {
        std::unique_lock<std::mutex> lk(qm);
        qcv.wait(lk, [&]() { return closed || !queue.empty(); });
        if (queue.empty()) return;
        job = std::move(queue.front());
        queue.pop_front();
      }
      qcv.notify_all();
      ERR_CODE rc = OK;
      bool handedOver = false;  // deliver() was called: the sink has moved on to the next block
      if (!opened) {
        rc = PROCESSING_ERROR;
        std::lock_guard<std::mutex> lk(qm);
        if (workerErr.empty()) workerErr = "no usable CUDA device (" + session.lastError() + "); this build has no CPU path";
      } else if (firstError.load() == OK) {  // (after a failure nothing more is written)
        zdwb_rows_out rows;
        memset(&rows, 0, sizeof(rows));
        // a regular output file: the rows stay on the device and go to their place in the file through the context's
        // pinned ring (no block-sized host buffer in between); a pipe / stdout: to the host, written in turn
        const bool direct = ordered.seekable();
        const int zrc = this->decodeBytes(session, job.bytes.data(), job.bytes.size(), true, job.firstRow, '\t', false, false, false,
                                          false, &rows, direct);
        if (zrc == ZDWB_OK && direct) {
          handedOver = true;
          uint64_t at = 0;
          bool ok = ordered.claim(job.seq, job.prefix.size() + (size_t)rows.len, &at);
          ok = ok && ordered.writePrefix(job.prefix, at) &&
               zdwb_device_to_fd(session.get(), rows.tsv, (size_t)rows.len, ordered.fd(), (long long)(at + job.prefix.size())) == ZDWB_OK;
          if (!ok) {
            ordered.fail();
            rc = FILE_CREATION_ERR;
          }
        } else if (zrc == ZDWB_OK) {
          handedOver = true;
          if (!ordered.deliver(job.seq, job.prefix, rows.tsv, rows.len)) rc = FILE_CREATION_ERR;
        } else {
          rc = zrc == ZDWB_ERR_CORRUPT ? CORRUPTED_DATA_ERROR
               : (zrc == ZDWB_ERR_ROW_COUNT || zrc == ZDWB_ERR_TRUNCATED) ? ROW_COUNT_ERR : PROCESSING_ERROR;
          std::lock_guard<std::mutex> lk(qm);
          if (workerErr.empty()) workerErr = zdwb_last_error(session.get());
        }
      }
      if (rc != OK) {  // (before the sink moves on: a later block that finds the sink failed must not report first)
        int expected = OK;
        firstError.compare_exchange_strong(expected, (int)rc);
      }
      if (!handedOver) ordered.skip(job.seq);
    }
  };
  vector<std::thread> threads;
  for (size_t w = 0; w < workers.size(); ++w) threads.push_back(std::thread(work, workers[w]));

  ERR_CODE rc = OK;
  size_t seq = 0;
  try {
    do {
      typename UnconvertFromZDW_Base::BlockInfo info;
      rc = this->peekBlock(info);
      if (rc != OK) break;
      this->rowsRead = 0;
      if (!this->bQuiet) this->statusOutput(INFO, "Reading %u rows\n", this->numLines);
      zdwb_rows_out sk;
      memset(&sk, 0, sizeof(sk));
      rc = this->decodeBlock(info, '\t', false, false, false, &sk, NULL, true);  // skim: where does the block end?
      if (rc != OK) break;
      if (firstError.load() != OK) break;
      DecodeJob job;
      job.seq = seq++;
      job.bytes.assign(this->input->data(), this->input->data() + (size_t)sk.consumed);
      job.firstRow = this->rowsBeforeBlock + 1;
      job.numLines = this->numLines;
      if (this->bOutputNonEmptyColumnHeader) job.prefix = this->getBlockHeaderString(info);
      {
        std::unique_lock<std::mutex> lk(qm);
        qcv.wait(lk, [&]() { return queue.size() < maxQueued; });
        queue.push_back(std::move(job));
      }
      qcv.notify_all();
      this->rowsRead = this->numLines;
      this->input->consume((size_t)sk.consumed);
      if (this->bShowStatus) this->statusOutput(INFO, "\r%u\n", this->rowsRead);
      this->rowsBeforeBlock += this->numLines;
      if (this->isLastBlock() && !this->bQuiet)
        this->statusOutput(INFO, "%s %s\n\n", displayName(this->inFileName).c_str(), "uncompressed");
    } while (!this->isLastBlock());
  } catch (...) {
    {
      std::lock_guard<std::mutex> lk(qm);
      closed = true;
    }
    qcv.notify_all();
    for (size_t w = 0; w < threads.size(); ++w) threads[w].join();
    ordered.finish();
    throw;
  }
  {
    std::lock_guard<std::mutex> lk(qm);
    closed = true;
  }
  qcv.notify_all();
  for (size_t w = 0; w < threads.size(); ++w) threads[w].join();
  const bool written = ordered.finish();
  if (rc == OK && firstError.load() != OK) {
    rc = (ERR_CODE)firstError.load();
    if (rc == ROW_COUNT_ERR) {
      this->printError(this->exeName, displayName(this->inFileName));
      this->statusOutput(INFO, "Rows unpacked (%u) does not match expected (%u)\n\n", 0u, this->numLines);
    } else if (rc == PROCESSING_ERROR) {
      this->statusOutput(ERROR, "%s: GPU decode failed: %s\n", this->exeName.empty() ? "UnconvertFromZDW" : this->exeName.c_str(),
                         workerErr.c_str());
    }
  }
  if (rc == OK && !written) rc = FILE_CREATION_ERR;
  return rc;
}

// Whole file to disk / stdout (reference :1656-1844).
template <typename BufferedOutput_T>
ERR_CODE UnconvertFromZDWToFile<BufferedOutput_T>::unconvert(const char* binaryName, const char* outputBasename, const char* ext,
                                                             const char* specifiedDir, bool bStdout) {
  if (binaryName && *binaryName) this->exeName = binaryName;
  if (!this->statusOutput) this->statusOutput = bStdout ? stdErrStatusOutputCallback : defaultStatusOutputCallback;
  if (!this->isReadOpen()) {
    this->statusOutput(ERROR, "%s: Could not open %s for reading\n", this->exeName.c_str(), displayName(this->inFileName).c_str());
    return FILE_OPEN_ERR;
  }
  string sourceDir, stub;
  if (!this->inFileName.empty()) {
    UnconvertFromZDW_Base::splitDirAndBase(this->inFileName, sourceDir, stub);
  } else {
    if (!outputBasename) bStdout = true;                       // stdin in, nothing named: stream out
    if (!specifiedDir || !*specifiedDir) specifiedDir = ".";
    stub = "stdin";
  }
  const string outputDir = (specifiedDir && *specifiedDir) ? string(specifiedDir) : sourceDir;
  const string outBase = outputBasename ? string(outputBasename) : stub;

  if (this->bShowStatus) {
    const char* what = this->bShowBasicStatisticsOnly ? "Showing statistics"
                       : this->bTestOnly ? "Testing"
                       : this->bOutputDescFileOnly ? "Outputting .desc file only"
                       : this->metadataOptions.bOutputOnlyMetadata ? "Outputting .metadata file only"
                       : "Processing";
    this->statusOutput(INFO, "\n%s %s\n", stub.c_str(), what);
  }

  ERR_CODE rc = this->readHeader();
  bool ownOut = false;
  if (rc != OK) {
    if (rc == UNSUPPORTED_ZDW_VERSION_ERR && this->version > UnconvertFromZDW_Base::UNCONVERT_ZDW_VERSION)
      this->statusOutput(ERROR, "%s: %s is newer (version %d) than supported version (%d)\n%s\n", this->exeName.c_str(), stub.c_str(),
                         this->version, UnconvertFromZDW_Base::UNCONVERT_ZDW_VERSION,
                         this->version > 10000 ? "Maybe you are trying to read a tar or gzip file?\n" : "");
    return rc;
  }

  const bool producesRows = !this->bTestOnly && !this->bShowBasicStatisticsOnly && !this->bOutputDescFileOnly &&
                            !this->metadataOptions.bOutputOnlyMetadata;
  if (producesRows) {
    const string outName = outputDir + "/" + outBase + (ext ? ext : "");
    if (this->bShowStatus) this->statusOutput(INFO, "Writing %s\n", outName.c_str());
    this->out = bStdout ? stdout : fopen(outName.c_str(), "w");
    ownOut = !bStdout;
    if (!this->out) {
      this->statusOutput(ERROR, "%s: Could not open %s for writing\n", this->exeName.c_str(), outName.c_str());
      return FILE_CREATION_ERR;
    }
  }
  struct Closer {
    FILE*& f;
    bool own;
    ~Closer() {
      if (f && own) fclose(f);
      f = NULL;
    }
  } closer = {this->out, ownOut};

  if (!this->bTestOnly && !this->bShowBasicStatisticsOnly) {
    if ((!bStdout || this->bOutputDescFileOnly) && !this->metadataOptions.bOutputOnlyMetadata) {
      const ERR_CODE d = bStdout ? this->outputDescToStdOut(this->columnNames)
                                 : this->outputDescToFile(this->columnNames, outputDir, outBase.c_str(), ext);
      if (d != OK) {
        this->statusOutput(ERROR, "%s: Could not extract the %s.desc%s file\n", this->exeName.c_str(), outBase.c_str(), ext ? ext : "");
        return d;
      }
      if (this->bOutputDescFileOnly) return OK;
    }
    if (!bStdout || this->metadataOptions.bOutputOnlyMetadata) {
      if (!this->metadata.empty() || !this->metadataOptions.keys.empty() || this->metadataOptions.bOutputOnlyMetadata) {
        const ERR_CODE m = bStdout ? this->outputMetadataToStdOut() : this->outputMetadataToFile(outputDir, outBase.c_str());
        if (m != OK) {
          this->statusOutput(ERROR, "%s: Could not extract the %s.metadata file\n", this->exeName.c_str(), outBase.c_str());
          return m;
        }
      }
      if (this->metadataOptions.bOutputOnlyMetadata) return OK;
    }
  }

  {
    BufferedOutput_T sink(this->out ? this->out : stdout);
    const size_t nWorkers = std::max<size_t>(1, this->gpuList.size()) * (size_t)std::max(1, this->lanesPerGpu);
    if (producesRows && nWorkers > 1) {
      rc = this->decodeBlocksFanOut(sink);
      if (rc != OK) return rc;
    } else {
      do {
        rc = this->parseNextBlock(sink);
        if (rc != OK) return rc;
      } while (!this->isLastBlock());
      if (!sink.waitIdle()) return FILE_CREATION_ERR;  // a deferred write came up short (disk full, closed pipe)
    }
  }
  if (!this->bShowBasicStatisticsOnly) {
    this->input->finalDummyRead();
    if (!this->isFinished()) {
      this->statusOutput(INFO, "Did not reach EOF\n");
      return ZDW_LONGER_THAN_EXPECTED_ERR;
    }
  }
  return OK;
}

template class UnconvertFromZDWToFile<BufferedOutput>;
template class UnconvertFromZDWToFile<BufferedOrderedOutput>;

// ---------------------------------------------------------------------------------------------------------------
// in-memory API (reference :1872-2142)
// ---------------------------------------------------------------------------------------------------------------
UnconvertFromZDWToMemory::UnconvertFromZDWToMemory(const string& fileName, const bool useInternalBuffer, const bool showStatus,
                                                   const bool quiet, const bool testOnly, const bool descOnly)
    : UnconvertFromZDW<BufferedOutputInMem>(fileName, showStatus, quiet, testOnly, descOnly), bUseInternalBuffer(useInternalBuffer),
      blockOpen(false), neededBufferSize(0), slab(NULL), slabRowOff(NULL), currentRowLength(0),
      bDecodeAhead(getenv("ZDW_NO_DECODE_AHEAD") == NULL), slabSession(1), aheadPending(false), aheadRc(-1) {
  statusOutput = defaultStatusOutputCallback;
  memset(&aheadRows, 0, sizeof(aheadRows));
}

UnconvertFromZDWToMemory::~UnconvertFromZDWToMemory() { joinDecodeAhead(); }

// The current block has just been decoded and its bytes consumed: `input` stands at the next block.  The helper buffers
// about as many bytes as the last block took and decodes them in the other context; what it finds is adopted by the next
// handleZDWParseBlockHeader if it is a complete block, and decoded again the ordinary way if not (a longer block).
void UnconvertFromZDWToMemory::startDecodeAhead() {
  if (!bDecodeAhead || aheadPending || isLastBlock()) return;
  const unsigned long long firstRow = rowsBeforeBlock + numLines + 1;
  const size_t want = lastBlockBytes ? lastBlockBytes + lastBlockBytes / 4 + 4096 : ((size_t)1 << 20);
  const int device = gpuList.empty() ? gpuDevice : gpuList[0];
  GpuSession* g = slabSession == 0 ? &gpu2 : &gpu;
  aheadPending = true;
  aheadRc = -1;
  aheadThread = std::thread([this, g, firstRow, want, device]() {
    try {
      input->ensure(want);
      if (input->available() < 10 || !g->open(device)) return;
      memset(&aheadRows, 0, sizeof(aheadRows));
      aheadRc = decodeBytes(*g, input->data(), input->available(), input->sourceEnded(), firstRow, '\0', true, false, false, false,
                            &aheadRows);
    } catch (...) {
      aheadRc = -1;  // (out of memory while buffering: the ordinary path reports it)
    }
  });
}

void UnconvertFromZDWToMemory::joinDecodeAhead() {
  if (aheadThread.joinable()) aheadThread.join();
}

// Decodes the next block in one go (NUL-separated fields, row offsets) and keeps it for getRow to hand out.
ERR_CODE UnconvertFromZDWToMemory::handleZDWParseBlockHeader() {
  joinDecodeAhead();
  const bool ahead = aheadPending && aheadRc == ZDWB_OK;
  aheadPending = false;
  BlockInfo info;
  ERR_CODE rc = peekBlock(info);
  if (rc != OK) return rc;
  rowsRead = 0;
  pendingHeaderLine.clear();
  if (bOutputNonEmptyColumnHeader) pendingHeaderLine = getBlockHeaderString(info);
  neededBufferSize = std::max((size_t)exportFileLineLength + virtualLineLength + 1,
                              pendingHeaderLine.empty() ? (size_t)0 : pendingHeaderLine.size() + 1);
  zdwb_rows_out rows;
  memset(&rows, 0, sizeof(rows));
  if (ahead && aheadRows.nrows == info.numLines) {
    rows = aheadRows;  // decoded while the caller walked the block before
    slabSession ^= 1;
    if (rows.consumed) lastBlockBytes = (size_t)rows.consumed;
  } else {
    // (the first block, a block longer than the helper buffered, or decode-ahead switched off)
    rc = decodeBlock(info, '\0', true, false, false, &rows, slabSession == 0 ? &gpu2 : &gpu);
    if (rc != OK) return rc;
    slabSession ^= 1;
  }
  slab = reinterpret_cast<const char*>(rows.tsv);
  slabRowOff = rows.row_off;
  input->consume((size_t)rows.consumed);
  blockOpen = true;
  setState(ZDW_OUTPUT_BLOCK_HEADER);
  startDecodeAhead();
  return OK;
}

// Hands one NUL-separated record to the caller: copied into *buffer (grown with new[] when too small, reference
// BufferedOutput.cpp:415-426) or, with the internal buffer, referenced in place.
ERR_CODE UnconvertFromZDWToMemory::deliver(const char* src, size_t len, size_t fields, char** buffer, size_t* size,
                                           const char** outColumns) {
  const char* base = src;
  if (!bUseInternalBuffer) {
    if (!buffer || !*buffer || !size) return BAD_PARAMETER;
    const size_t need = std::max(neededBufferSize, len);
    if (need > *size) {
      delete[] *buffer;
      *buffer = new char[need];
      *size = need;
    }
    memcpy(*buffer, src, len);
    base = *buffer;
  }
  currentRowLength = len ? len - 1 : 0;
  if (outColumns) {
    const char* at = base;
    for (size_t k = 0; k < fields; ++k) {
      outColumns[k] = at;
      at += strlen(at) + 1;
    }
  }
  return OK;
}

ERR_CODE UnconvertFromZDWToMemory::getRow(const char** outColumns) {
  size_t n;
  return getRow(NULL, NULL, outColumns, n);
}

ERR_CODE UnconvertFromZDWToMemory::getRow(char** buffer, size_t* size, const char** outColumns, size_t& numCols) {
  for (;;) {
    switch (eState) {
      case ZDW_BEGIN: {
        const ERR_CODE rc = readHeader();
        if (rc != OK) return rc;
        break;
      }
      case ZDW_PARSE_BLOCK_HEADER: {
        const ERR_CODE rc = handleZDWParseBlockHeader();
        if (rc != OK) return rc;
        break;
      }
      case ZDW_OUTPUT_BLOCK_HEADER:
        setState(ZDW_GET_NEXT_ROW);
        if (!pendingHeaderLine.empty()) {
          // the block header line is returned as a row of its own (raw line, one "column")
          if (bUseInternalBuffer) {
            internalRow.assign(pendingHeaderLine.begin(), pendingHeaderLine.end());
            internalRow.push_back('\0');
            currentRowLength = pendingHeaderLine.size();
            if (outColumns) outColumns[0] = internalRow.data();
          } else {
            const string line = pendingHeaderLine + '\0';
            const ERR_CODE rc = deliver(line.data(), line.size(), 1, buffer, size, outColumns);
            if (rc != OK) return rc;
          }
          pendingHeaderLine.clear();
          return OK;
        }
        break;
      case ZDW_GET_NEXT_ROW:
        if (rowsRead < numLines) {
          const uint64_t a = slabRowOff[rowsRead], b = slabRowOff[rowsRead + 1];
          ++rowsRead;
          const ERR_CODE rc = deliver(slab + a, (size_t)(b - a), numOutputColumns(), buffer, size, outColumns);
          numCols = numOutputColumns();
          return rc;
        }
        blockOpen = false;
        rowsBeforeBlock += numLines;
        setState(isLastBlock() ? ZDW_FINISHING : ZDW_PARSE_BLOCK_HEADER);
        break;
      case ZDW_FINISHING:
        input->finalDummyRead();
        setState(ZDW_END);
        return isFinished() ? AT_END_OF_FILE : ZDW_LONGER_THAN_EXPECTED_ERR;
      case ZDW_END:
        return AT_END_OF_FILE;
    }
  }
}

ERR_CODE UnconvertFromZDWToMemory::getNumOutputColumns(size_t& num) {
  for (;;) {
    switch (eState) {
      case ZDW_BEGIN: {
        const ERR_CODE rc = readHeader();
        if (rc != OK) return rc;
        break;
      }
      case ZDW_PARSE_BLOCK_HEADER: {
        const ERR_CODE rc = handleZDWParseBlockHeader();
        if (rc != OK) return rc;
        break;
      }
      case ZDW_FINISHING:
        return UNSUPPORTED_OPERATION;
      default:
        if (!blockOpen) return PROCESSING_ERROR;
        num = numOutputColumns();
        return OK;
    }
  }
}

size_t UnconvertFromZDWToMemory::getCurrentRowLength() { return currentRowLength; }

void UnconvertFromZDWToMemory::getColumnNamesVector(vector<string>& out) {
  map<int, string> ordered;
  for (size_t c = 0; c < columnNames.size(); ++c) {
    const int at = namesOfColumnsToOutput.empty() ? (int)c : outputColumns[c];
    if (at != IGNORE_COLUMN) ordered[at] = columnNames[c];
  }
  for (map<int, string>::const_iterator it = blankColumnNames.begin(); it != blankColumnNames.end(); ++it)
    ordered[it->first] = it->second;
  for (map<int, string>::const_iterator it = ordered.begin(); it != ordered.end(); ++it) out.push_back(it->second);
}

bool UnconvertFromZDWToMemory::hasColumnName(const string& name) const {
  return std::find(columnNames.begin(), columnNames.end(), name) != columnNames.end();
}

bool UnconvertFromZDWToMemory::OutputDescToFile(const string& outputDir) {
  string dir, base;
  splitDirAndBase(inFileName, dir, base);
  return outputDescToFile(columnNames, outputDir, base.c_str(), ".sql") == OK;
}

// metadata key "lineage" = base,rows|base,rows|... (reference :2103-2142)
vector<std::pair<uint64_t, string> > UnconvertFromZDWToMemory::getFileLineage() {
  vector<std::pair<uint64_t, string> > out;
  if (eState == ZDW_BEGIN && readHeader() != OK) {
    out.push_back(std::make_pair((uint64_t)0, string("bad ZDW header")));
    return out;
  }
  map<string, string>::const_iterator it = metadata.find("lineage");
  if (it == metadata.end() || it->second.empty()) return out;
  const string& v = it->second;
  size_t at = 0;
  while (at != string::npos) {
    const size_t comma = v.find(',', at);
    if (comma == string::npos) {
      out.push_back(std::make_pair((uint64_t)0, string("bad lineage data")));
      return out;
    }
    out.push_back(std::make_pair((uint64_t)strtoull(v.c_str() + comma + 1, NULL, 10), v.substr(at, comma - at)));
    const size_t bar = v.find('|', comma + 1);
    at = bar == string::npos ? string::npos : bar + 1;
  }
  return out;
}

}  // namespace zdw
}  // namespace adobe
