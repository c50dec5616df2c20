// convertDWfile -- command line front end of the TSV -> ZDW encoder (B200 build).
// Flags, messages and exit codes follow the reference CLI (cplusplus/convertDWfile.cpp:44-66, :98-260); the
// --rows-per-block / --block-plan / --block-bytes / --gpu options are additions of this build.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <map>
#include <string>
#include <vector>

#include "ConvertToZDW.h"

using adobe::zdw::ConvertToZDW;

namespace {

const char* baseName(const char* path) {
  const char* s = strrchr(path, '/');
  return s ? s + 1 : path;
}

void printVersion() {
  printf("ConvertToZDW, Version %i%s\n", ConvertToZDW::CONVERT_ZDW_CURRENT_VERSION, ConvertToZDW::CONVERT_ZDW_VERSION_TAIL);
}

void printUsage(const char* exe) {
  printf("Usage: %s [-d <dir>] [-(b|J|Jf|z|q|r|v)] [other options] file1 [file2] ...\n", baseName(exe));
  fputs("\t-b  compress .zdw with bzip2 [default=use gzip]\n"
        "\t-J  compress .zdw with xz [default=use gzip]\n"
        "\t -Jf  compress to .zdw.xz file via fsx (applying fastlzma2 algorithm)\n"
        "\t-z  compress .zdw with zstd\n"
        "\t-d  output to directory <dir> [default=same directory as source file]\n"
        "\t-i  streaming input from stdin; file1 is used as the implied name for the input stream\n"
        "\t-q  quiet operation (no status or progress messages) [default=not quiet]\n"
        "\t-r  remove the old files\n"
        "\t-t  trim trailing spaces from fields (for MySQL 5 exports)\n"
        "\t-v  validate the new file\n"
        "\n"
        "\t--zargs=X          arguments to pass in to the file compression process\n"
        "\t--mem-limit=<MB>   limit the MB of RAM used (default=3072 MB)\n"
        "\n"
        "\t--metadata:<key>=<value>   supply a key-value pair to store as file metadata for every file being converted\n"
        "\t--metadata-file=<filename> supply a filepath to specify key-value pairs (formatted as '<key>=<value>' pairs, each on a separate line) to store as file metadata for every file being converted\n"
        "\n"
        "\t--rows-per-block=<N>  (B200 build) close a ZDW block every N rows [default=no row limit]\n"
        "\t--block-plan=<rows:spill,...>  (B200 build) explicit blocks: rows per block and the columns of the next row that\n"
        "\t                      (B200 build) were already parsed when the reference closed the block (reproduces its memory-driven cuts)\n"
        "\t--block-bytes=<N>     (B200 build) TSV bytes per block window [default=1 GiB]\n"
        "\t--heap-blocks=<K>     (B200 build) close every block where the reference does when it finds itself over --mem-limit at\n"
        "\t                      (B200 build) its K-th 64 MiB string-heap allocation; --mem-limit itself is accepted and has no effect\n"
        "\t--gpu=<N>             (B200 build) CUDA device to use [default=$ZDW_GPU or 0]\n"
        "\t--gpus=<N|all|a,b,..> (B200 build) spread the blocks of a file over N GPUs / all GPUs / the listed devices\n"
        "\t                      (B200 build) [default=$ZDW_GPUS or the one device]; the output does not depend on it\n"
        "\t--lanes-per-gpu=<N>   (B200 build) encode workers (context + pinned window + host thread) per GPU [default=2]\n"
        "\n"
        "\t--help     show this help\n"
        "\t--version  show the version number\n"
        "Input files must have a .sql extension.\n"
        "\n",
        stdout);
}

int reportFailure(ConvertToZDW::ERR_CODE code) {
  const int idx = code < ConvertToZDW::ERR_CODE_COUNT ? code : ConvertToZDW::UNKNOWN_ERROR;
  fprintf(stderr, "ZDW conversion failed.  Internal error code=%i (%s)\n", code, ConvertToZDW::ERR_CODE_TEXTS[idx]);
  return code;
}

int unknownParameter(const char* exe, const char* arg) {
  fprintf(stderr, "%s: Unknown parameter '%s'\n\n", exe, arg);
  fprintf(stderr, "    Run with --help for usage info.\n");
  return ConvertToZDW::BAD_PARAMETER;
}

struct Options {
  bool streaming, removeOld, trim, validate, quiet;
  ConvertToZDW::Compressor compressor;
  const char* outputDir;
  const char* zArgs;
  std::map<std::string, std::string> metadata;
  unsigned long long rowsPerBlock;
  std::vector<std::pair<uint64_t, uint32_t> > blockPlan;
  unsigned long long blockBytes;
  int gpu;
  std::string gpus;
  int lanes;
  unsigned heapBlocks;
  bool memLimitGiven;
  std::vector<const char*> files;
  Options()
      : streaming(false), removeOld(false), trim(false), validate(false), quiet(false), compressor(ConvertToZDW::GZIP),
        outputDir(NULL), zArgs(NULL), rowsPerBlock(0), blockBytes(0), gpu(-1), lanes(0), heapBlocks(0), memLimitGiven(false) {}
};

}  // namespace

int main(int argc, char* argv[]) {
  const char* exe = argv[0];
  if (argc < 2) {
    printVersion();
    printUsage(exe);
    return ConvertToZDW::NO_ARGS;
  }
  Options opt;
  for (int i = 1; i < argc; ++i) {
    const char* a = argv[i];
    if (a[0] != '-') {
      if (opt.streaming && !opt.files.empty()) return reportFailure(ConvertToZDW::TOO_MANY_INPUT_FILES);
      opt.files.push_back(a);
      continue;
    }
    switch (a[1]) {
      case 'b': opt.compressor = ConvertToZDW::BZIP2; break;
      case 'J': opt.compressor = a[2] == 'f' ? ConvertToZDW::FXZ : ConvertToZDW::XZ; break;
      case 'z': opt.compressor = ConvertToZDW::ZSTD; break;
      case 'd':
        if (++i >= argc) {
          printUsage(exe);
          return ConvertToZDW::MISSING_ARGUMENT;
        }
        opt.outputDir = argv[i];
        break;
      case 'i': opt.streaming = true; break;
      case 'q': opt.quiet = true; break;
      case 'r': opt.removeOld = true; break;
      case 't': opt.trim = true; break;
      case 'v': opt.validate = true; break;
      case '-': {
        const char* flag = a + 2;
        if (!strcmp(flag, "help")) {
          printVersion();
          printUsage(exe);
          return ConvertToZDW::OK;
        }
        if (!strcmp(flag, "ver") || !strcmp(flag, "version")) {
          printVersion();
          return ConvertToZDW::OK;
        }
        if (!strncmp(flag, "mem-limit=", 10)) {
          // The reference compares its process's virtual memory with this limit to decide when a block ends
          // (memory.cpp:65-81, stringheap.cpp:75-86): a property of that process, not of the input.  The flag is
          // accepted and has no effect here; --heap-blocks=K reproduces the cut itself (the K-th 64 MiB string-heap
          // allocation), --block-bytes / --rows-per-block set this build's own block size.
          const double mb = atof(flag + 10);
          if (!(mb > 0.0)) return unknownParameter(exe, a);
          opt.memLimitGiven = true;
          break;
        }
        if (!strncmp(flag, "heap-blocks=", 12)) {
          opt.heapBlocks = (unsigned)strtoul(flag + 12, NULL, 10);
          if (!opt.heapBlocks) return reportFailure(ConvertToZDW::BAD_PARAMETER);
          break;
        }
        if (!strncmp(flag, "metadata:", 9)) {
          const char* key = flag + 9;
          const char* eq = strchr(key, '=');
          if (!eq) return unknownParameter(exe, a);
          opt.metadata[std::string(key, eq - key)] = std::string(eq + 1);
          break;
        }
        if (!strncmp(flag, "metadata-file=", 14)) {
          const int line = ConvertToZDW::loadMetadataFile(flag + 14, opt.metadata);
          if (line) {
            fprintf(stderr, "%s: Metadata file load error '%s' (line %d)\n\n", exe, flag + 14, line);
            return ConvertToZDW::BAD_PARAMETER;
          }
          break;
        }
        if (!strncmp(flag, "zargs=", 6)) {
          opt.zArgs = flag + 6;
          break;
        }
        if (!strncmp(flag, "rows-per-block=", 15)) {
          opt.rowsPerBlock = strtoull(flag + 15, NULL, 10);
          break;
        }
        if (!strncmp(flag, "block-plan=", 11)) {  // rows:spill,rows:spill,...
          const char* at = flag + 11;
          while (*at) {
            char* end = NULL;
            const unsigned long long rows = strtoull(at, &end, 10);
            unsigned long spill = 0;
            if (end && *end == ':') spill = strtoul(end + 1, &end, 10);
            if (!end || end == at || rows == 0) return reportFailure(ConvertToZDW::BAD_PARAMETER);
            opt.blockPlan.push_back(std::make_pair((uint64_t)rows, (uint32_t)spill));
            at = *end == ',' ? end + 1 : end;
            if (*end && *end != ',') return reportFailure(ConvertToZDW::BAD_PARAMETER);
          }
          break;
        }
        if (!strncmp(flag, "block-bytes=", 12)) {
          opt.blockBytes = strtoull(flag + 12, NULL, 10);
          break;
        }
        if (!strncmp(flag, "gpu=", 4)) {
          opt.gpu = atoi(flag + 4);
          break;
        }
        if (!strncmp(flag, "gpus=", 5)) {
          opt.gpus = flag + 5;
          break;
        }
        if (!strncmp(flag, "lanes-per-gpu=", 14)) {
          opt.lanes = atoi(flag + 14);
          if (opt.lanes < 1) return reportFailure(ConvertToZDW::BAD_PARAMETER);
          break;
        }
        printUsage(exe);
        return unknownParameter(exe, a);
      }
      default:
        return unknownParameter(exe, a);
    }
  }
  if (opt.files.empty()) return reportFailure(ConvertToZDW::NO_INPUT_FILES);
  if (opt.streaming && isatty(0)) return reportFailure(ConvertToZDW::NO_INPUT_FILES);  // nothing is piped in

  // --gpus: a count, "all", or a list of devices (a device may be named more than once)
  std::vector<int> gpuList;
  {
    std::string spec = opt.gpus;
    if (spec.empty() && getenv("ZDW_GPUS")) spec = getenv("ZDW_GPUS");
    if (!adobe::zdw::parseGpuSpec(spec, opt.gpu, gpuList)) return reportFailure(ConvertToZDW::BAD_PARAMETER);
  }

  int exitCode = ConvertToZDW::OK;
  for (size_t f = 0; f < opt.files.size(); ++f) {
    if (f + 1 == opt.files.size()) adobe::zdw::GpuSession::processExiting() = true;
    std::vector<char> stub(strlen(opt.files[f]) + 1024);
    ConvertToZDW conv(opt.quiet, opt.streaming);
    conv.compressor = opt.compressor;
    if (opt.trim) conv.trimTrailingSpaces();
    if (opt.rowsPerBlock) conv.setRowsPerBlock(opt.rowsPerBlock);
    if (!opt.blockPlan.empty()) conv.setBlockPlan(opt.blockPlan);
    if (opt.blockBytes) conv.setBlockBytes((size_t)opt.blockBytes);
    if (opt.heapBlocks) conv.setHeapBlocks(opt.heapBlocks);
    if (opt.memLimitGiven && !opt.quiet && f == 0)
      fprintf(stderr, "%s: --mem-limit has no effect in this build (blocks are cut by --block-bytes, --rows-per-block or --heap-blocks)\n", exe);
    conv.setGpuDevice(opt.gpu);
    if (!gpuList.empty()) conv.setGpus(gpuList);
    if (opt.lanes) conv.setLanesPerGpu(opt.lanes);
    const ConvertToZDW::ERR_CODE res =
      conv.convertFile(opt.files[f], exe, opt.validate, stub.data(), opt.outputDir, opt.zArgs, opt.metadata);
    if (res != ConvertToZDW::OK) {
      if (!opt.quiet) reportFailure(res);
      exitCode = ConvertToZDW::CONVERSION_FAILED;  // every per-file failure exits with 2 (:230-237)
    }
    if (opt.removeOld) {
      if (res != ConvertToZDW::OK) {
        fprintf(stderr, "Could not remove original %s file because conversion was not good\n", stub.data());
      } else {
        const std::string base = stub.data();
        unlink((base + ".desc." + conv.getInputFileExtension()).c_str());
        unlink((base + "." + conv.getInputFileExtension()).c_str());
      }
    }
  }
  fflush(NULL);
  _exit(exitCode);  // (no CUDA teardown: see GpuSession::processExiting)
}
