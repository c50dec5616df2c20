// api_rowloop -- TEST INFRASTRUCTURE: times the row-at-a-time decode API (UnconvertFromZDWToMemory::getRow, the loop of
// the reference's cplusplus/test_unconvert_api.cpp:97-116 without the printf) and prints one JSON line.  The same
// source is compiled twice: against the unmodified reference (oracle/Makefile -> oracle/_ref/api_rowloop) and against
// this repository's host classes (zdw_b200/host/Makefile -> zdw_b200/bin/api_rowloop), so the two numbers measure the
// same calls.  A checksum over every field byte makes sure both sides really produced the same rows.
#include <stdio.h>
#include <string.h>
#include <time.h>

#include <string>

#include "zdw/UnconvertFromZDW.h"

using namespace adobe::zdw;

static double now_s() {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

int main(int argc, char* argv[]) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s [--checksum] file.zdw\n", argv[0]);
    return 1;
  }
  // --checksum: FNV-1a over every field byte (parity runs); without it only the field lengths are taken, so that the
  // timing is the API's and not the checksum's
  const bool checksum = argc > 2 && !strcmp(argv[1], "--checksum");
  const double t0 = now_s();
  UnconvertFromZDWToMemory dec(argv[checksum ? 2 : 1], false);
  ERR_CODE rc = dec.readHeader();
  if (rc != OK) {
    fprintf(stderr, "Error %i\n", rc);
    return rc;
  }
  size_t numColumns = 0;
  rc = dec.getNumOutputColumns(numColumns);
  if (rc != OK) {
    fprintf(stderr, "Error %i\n", rc);
    return rc;
  }
  const char** columns = new const char*[numColumns];
  size_t lineLength = dec.getLineLength();
  char* buffer = new char[lineLength];
  unsigned long long rows = 0, bytes = 0, sum = 1469598103934665603ull;
  const double t1 = now_s();
  while (!dec.isFinished()) {
    rc = dec.getRow(&buffer, &lineLength, columns, numColumns);
    if (rc == OK) {
      ++rows;
      for (size_t c = 0; c < numColumns; ++c) {
        const char* s = columns[c];
        const size_t n = strlen(s);
        bytes += n + 1;  // the field and its separator / line end
        if (checksum) {
          for (size_t k = 0; k < n; ++k) sum = (sum ^ (unsigned char)s[k]) * 1099511628211ull;
          sum = (sum ^ 0xffu) * 1099511628211ull;
        }
      }
    } else if (rc != AT_END_OF_FILE) {
      fprintf(stderr, "Error %i\n", rc);
      return rc;
    }
  }
  const double t2 = now_s();
  printf("{\"rows\": %llu, \"tsv_bytes\": %llu, \"fnv1a\": \"%016llx\", \"open_s\": %.6f, \"loop_s\": %.6f, \"mb_per_s\": %.1f, "
         "\"loop_mb_per_s\": %.1f}\n",
         rows, bytes, checksum ? sum : 0ull, t1 - t0, t2 - t1, (double)bytes / 1e6 / (t2 - t0), (double)bytes / 1e6 / (t2 - t1));
  delete[] buffer;
  delete[] columns;
  return 0;
}
