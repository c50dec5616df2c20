// unconvertDWfile -- command line front end of the ZDW -> TSV decoder (B200 build).
// Flags, validation order, messages and exit codes follow the reference CLI (cplusplus/unconvertDWfile.cpp:44-80,
// :207-355): the exit code is the decoder's ERR_CODE.  --gpu=N is an addition of this build.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <string>

#include "zdw/UnconvertFromZDW.h"

using namespace adobe::zdw;
using std::string;

namespace {

int g_gpu = -1;
int g_lanes = 0;
std::string g_gpusSpec;
std::vector<int> g_gpuList;

const char* baseName(const char* path) {
  const char* s = strrchr(path, '/');
  return s ? s + 1 : path;
}

void printVersion() {
  printf("UnconvertFromZDW, Version %d%s\n", UnconvertFromZDW_Base::UNCONVERT_ZDW_VERSION, UnconvertFromZDW_Base::UNCONVERT_ZDW_VERSION_TAIL);
}

void printUsage(const char* exe) {
  printf("Usage: %s [-(i|o|q|s|t|v|w)] [-c[e|i|x] csvColumnNames] [other options] file1 [file2...]\n", baseName(exe));
  fputs("\t-  direct outputted text to stdout, and status text to stderr\n"
        "\t     No .desc file is outputted, except when the -o option is also set.\n"
        "\t-a <text to append>  specify text to be appended to the output filename\n"
        "\t-c specify a comma-separated list of column names to output (default = all columns)\n"
        "\t\t Columns are output in the order they are given.\n"
        "\t\t Non-existent and duplicate column names result in an error.\n"
        "\t-ce same as '-c', but provide an empty text column\n"
        "\t\t when a requested column is not present.\n"
        "\t-ci same as '-c', but do not error when invalid columns are specified\n"
        "\t\t Non-existent and duplicate column names after the first entry are ignored.\n"
        "\t-cx Include all columns except for this comma-separated list\n"
        "\t-d <outputDirectory>  specify the directory in which to place the resulting files\n"
        "\t\t (default=the files will be placed in the same directory as the .zdw file)\n"
        "\t-i read data to unconvert from stdin and by default send it to stdout.\n"
        "\t     No filenames listed on the command line will be processed.\n"
        "\t     If a filename is specified, this will be used as the output filename.\n"
        "\t-o write the .desc file to disk (or stdout) and then exit\n"
        "\t-q quiet -- no progress output (overrides -v)\n"
        "\t-s show basic file statistics only\n"
        "\t-t test integrity of zdw file only\n"
        "\t-v verbose -- show count of rows during conversion\n"
        "\t-w give outputted files no extension (default = .sql)\n"
        "\n"
        "\t--metadata       Provide to have only the metadata artifact output\n"
        "\t--metadata-keys  Provide to have only the metadata keys output (no values)\n"
        "\t--metadata-values=<csv keynames>  If supplied, only the indicated key-value pairs will be output\n"
        "\t\t Non-existent and duplicate keys result in an error.\n"
        "\t--metadata-values-allow-missing=<csv keynames>  If supplied, only the indicated key-value pairs will be output\n"
        "\t\t For keys not present in the file, an empty value will be supplied.\n"
        "\t\t This option is not compatible with --metadata-values\n"
        "\n"
        "\t--non-empty-column-header   output a header line listing non-empty columns in the next file block\n"
        "\t--gpu=<N>   (B200 build) CUDA device to use [default=$ZDW_GPU or 0]\n"
        "\t--gpus=<N|all|a,b,..>  (B200 build) spread the blocks over N GPUs / all GPUs / the listed devices [default=$ZDW_GPUS or one]\n"
        "\t--lanes-per-gpu=<N>    (B200 build) decode workers per GPU [default=2]; the output does not depend on either\n"
        "\n"
        "\t--help     show this help\n"
        "\t--version  show the version number\n"
        "\n",
        stdout);
}

int complain(const char* exe, const char* what, const char* arg) {
  fprintf(stderr, "%s: %s '%s'%s\n\n", exe, what, arg, "");
  fprintf(stderr, "    Run with --help for usage info.\n");
  return BAD_PARAMETER;
}
int unknownParameter(const char* exe, const char* arg) { return complain(exe, "Unknown parameter", arg); }
int missingArgument(const char* exe, const char* arg) {
  fprintf(stderr, "%s: Missing argument after parameter '%s'\n\n", exe, arg);
  fprintf(stderr, "    Run with --help for usage info.\n");
  return BAD_PARAMETER;
}
int extraOption(const char* exe, const char* arg) {
  fprintf(stderr, "%s: Extra option '%s' not allowed in tandem with other mutually exclusive options.\n\n", exe, arg);
  fprintf(stderr, "    Run with --help for usage info.\n");
  return BAD_PARAMETER;
}

struct Job {
  string ext;                 // what follows the base name of every output file
  string columns;             // -c* list
  COLUMN_INCLUSION_RULE rule;
  string dir;
  bool showStatus, quiet, testOnly, descOnly, toStdout, statsOnly, blockHeaders;
  internal::MetadataOptions meta;
  Job() : rule(FAIL_ON_INVALID_COLUMN), showStatus(false), quiet(false), testOnly(false), descOnly(false), toStdout(false),
          statsOnly(false), blockHeaders(false) {}
};

template <class Decoder>
ERR_CODE drive(Decoder& d, const Job& job, const char* exe, const char* outputBasename, const string& ext) {
  d.setGpuDevice(g_gpu);
  if (!g_gpuList.empty()) d.setGpus(g_gpuList);
  if (g_lanes) d.setLanesPerGpu(g_lanes);
  d.setMetadataOptions(job.meta);
  if (job.statsOnly) d.showBasicStatisticsOnly();
  d.outputNonEmptyColumnHeader(job.blockHeaders);
  return d.unconvert(exe, outputBasename, ext.c_str(), job.dir.c_str(), job.toStdout);
}

// One input (a file, or stdin when `file` is empty): picks the output policy like the reference (:143-160).
ERR_CODE unconvertOne(const string& file, const Job& job, const char* exe, const char* outputBasename, const string& ext) {
  ERR_CODE rc;
  if (job.columns.empty() || job.statsOnly) {
    UnconvertFromZDWToFile<BufferedOutput> d(file, job.showStatus, job.quiet, job.testOnly, job.descOnly);
    rc = drive(d, job, exe, outputBasename, ext);
  } else {
    UnconvertFromZDWToFile<BufferedOrderedOutput> d(file, job.showStatus, job.quiet, job.testOnly, job.descOnly);
    if (!d.setNamesOfColumnsToOutput(job.columns, job.rule)) rc = BAD_REQUESTED_COLUMN;
    else rc = drive(d, job, exe, outputBasename, ext);
  }
  if (rc != OK) {
    if (rc == NO_COLUMNS_TO_OUTPUT && job.descOnly) return OK;  // schema-only runs may select nothing
    fprintf(stderr, "Error code=%d (%s): ", rc, UnconvertFromZDW_Base::ERR_CODE_TEXTS[rc < ERR_CODE_COUNT ? rc : ERR_CODE_COUNT]);
    fprintf(stderr, "%s: %s failed\n\n", exe, !file.empty() ? file.c_str() : "from stdin");
  }
  return rc;
}

}  // namespace

// leaves without the CUDA teardown once the contexts were left standing (GpuSession::processExiting)
static int finish(int code) {
  if (!adobe::zdw::GpuSession::processExiting()) return code;
  fflush(NULL);
  _exit(code);
}

int main(int argc, char* argv[]) {
  const char* exe = argv[0];
  Job job;
  const char* appendText = NULL;
  string defaultExt = ".sql";
  bool fromStdin = false, sawColumns = false;
  if (argc < 2) {
    printVersion();
    printUsage(exe);
  }

  // pass 1: every argument is validated before any file is touched
  for (int i = 1; i < argc; ++i) {
    const char* a = argv[i];
    if (a[0] != '-') continue;
    const char f = a[1];
    if (strchr("aioqstvw", f) && f != '\0' && a[2] != '\0') return unknownParameter(exe, a);
    if (f == 'c') {
      if (sawColumns) return extraOption(exe, a);
      sawColumns = true;
      if (a[2] != '\0' && (!strchr("eix", a[2]) || a[3] != '\0')) return unknownParameter(exe, a);
    }
    if ((f == 'a' || f == 'c' || f == 'd') && i + 1 >= argc) return missingArgument(exe, a);
    switch (f) {
      case '\0': job.toStdout = true; break;
      case 'a': appendText = argv[++i]; break;
      case 'c':
        job.rule = a[2] == 'e' ? PROVIDE_EMPTY_MISSING_COLUMNS : a[2] == 'i' ? SKIP_INVALID_COLUMN
                   : a[2] == 'x' ? EXCLUDE_SPECIFIED_COLUMNS : FAIL_ON_INVALID_COLUMN;
        job.columns = argv[++i];
        break;
      case 'd':
        job.dir = argv[++i];
        if (!job.dir.empty() && job.dir[job.dir.size() - 1] == '/') job.dir.resize(job.dir.size() - 1);
        break;
      case 'i': fromStdin = true; break;
      case 'o': job.descOnly = true; break;
      case 'q': job.quiet = true; break;
      case 's': job.statsOnly = true; break;
      case 't': job.testOnly = true; break;
      case 'v': job.showStatus = true; break;
      case 'w': defaultExt.clear(); break;
      case '-': {
        const char* flag = a + 2;
        if (!strcmp(flag, "help")) {
          printVersion();
          printUsage(exe);
          return OK;
        }
        if (!strcmp(flag, "ver") || !strcmp(flag, "version")) {
          printVersion();
          return OK;
        }
        if (!strcmp(flag, "non-empty-column-header")) {
          job.blockHeaders = true;
          break;
        }
        if (!strcmp(flag, "metadata")) {
          job.meta.bOutputOnlyMetadata = true;
          break;
        }
        if (!strcmp(flag, "metadata-keys")) {
          job.meta.bOnlyMetadataKeys = true;
          break;
        }
        if (!strncmp(flag, "gpu=", 4)) {
          g_gpu = atoi(flag + 4);
          break;
        }
        if (!strncmp(flag, "gpus=", 5)) {
          g_gpusSpec = flag + 5;
          break;
        }
        if (!strncmp(flag, "lanes-per-gpu=", 14)) {
          g_lanes = atoi(flag + 14);
          if (g_lanes < 1) g_lanes = 1;
          break;
        }
        const bool strict = !strncmp(flag, "metadata-values=", 16);
        const bool lenient = !strncmp(flag, "metadata-values-allow-missing=", 30);
        if (strict || lenient) {
          if (!job.meta.keys.empty()) return extraOption(exe, a);
          job.meta.bAllowMissingKeys = lenient;
          string list = flag + (lenient ? 30 : 16);
          size_t at = 0;
          for (;;) {
            const size_t comma = list.find(',', at);
            const string key = list.substr(at, comma == string::npos ? string::npos : comma - at);
            if (!job.meta.keys.insert(key).second) return unknownParameter(exe, a);  // duplicate key
            if (comma == string::npos) break;
            at = comma + 1;
          }
          break;
        }
        return unknownParameter(exe, a);
      }
      default:
        return unknownParameter(exe, a);
    }
  }
  if (job.descOnly && job.meta.bOutputOnlyMetadata) {
    fprintf(stderr, "-o and --metadata options are incompatible.  Aborting.\n");
    return BAD_PARAMETER;
  }

  {
    std::string spec = g_gpusSpec;
    if (spec.empty() && getenv("ZDW_GPUS")) spec = getenv("ZDW_GPUS");
    if (!adobe::zdw::parseGpuSpec(spec, g_gpu, g_gpuList)) {
      fprintf(stderr, "%s: bad --gpus value '%s'\n", exe, spec.c_str());
      return BAD_PARAMETER;
    }
  }

  // pass 2: the files
  const char* stdinOutputName = NULL;
  int lastFileArg = 0;  // the last file is decoded without tearing its CUDA contexts down (the process exits next)
  for (int i = 1; i < argc; ++i) {
    if (argv[i][0] == '-') {
      if (argv[i][1] == 'a' || argv[i][1] == 'c' || argv[i][1] == 'd') ++i;
      continue;
    }
    lastFileArg = i;
  }
  for (int i = 1; i < argc; ++i) {
    const char* a = argv[i];
    if (a[0] == '-') {
      if (a[1] == 'a' || a[1] == 'c' || a[1] == 'd') ++i;
      continue;
    }
    if (i == lastFileArg && !fromStdin) adobe::zdw::GpuSession::processExiting() = true;
    if (a[0] == '\0') {
      fprintf(stderr, "%s: Empty filename not allowed\n\n", exe);
      fprintf(stderr, "    Run with --help for usage info.\n");
      return BAD_PARAMETER;
    }
    if (fromStdin) {
      stdinOutputName = a;  // with -i a name on the command line names the output
      continue;
    }
    const ERR_CODE rc = unconvertOne(a, job, exe, NULL, defaultExt + (appendText ? appendText : ""));
    if (rc != OK) return finish(rc);
  }
  if (fromStdin) {
    adobe::zdw::GpuSession::processExiting() = true;
    const ERR_CODE rc = unconvertOne("", job, exe, stdinOutputName, appendText ? string(appendText) : defaultExt);
    if (rc != OK) return finish(rc);
  }
  return finish(OK);
}
