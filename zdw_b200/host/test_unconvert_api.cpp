// test_unconvert_api -- drives the row-at-a-time decode API exactly the way the reference's example does
// (cplusplus/test_unconvert_api.cpp:54-129): caller-owned row buffer, one getRow per row, rows printed tab-joined.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "zdw/UnconvertFromZDW.h"

using namespace adobe::zdw;

static void usage(const char* exe) {
  printf("UnconvertFromZDWToMemory API test, Version %s\n", UnconvertFromZDW_Base::getVersion().c_str());
  printf("Usage: %s [-ci csvColumnNames] file1 [file2...]\n", exe);
}

static void printRow(const char** columns, size_t n) {
  for (size_t c = 0; c < n; ++c) {
    fputs(columns[c], stdout);
    if (c + 1 < n) putchar('\t');
  }
  putchar('\n');
}

int main(int argc, char* argv[]) {
  if (argc < 2) {
    usage(argv[0]);
    return 1;
  }
  int first = 1;
  std::string wanted;
  if (!strcmp(argv[1], "-ci")) {
    if (argc < 4) {
      usage(argv[0]);
      return 1;
    }
    wanted = argv[2];
    first = 3;
  }
  for (int i = first; i < argc; ++i) {
    UnconvertFromZDWToMemory dec(argv[i], false);
    if (!wanted.empty()) dec.setNamesOfColumnsToOutput(wanted, SKIP_INVALID_COLUMN);
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
    while (!dec.isFinished()) {
      rc = dec.getRow(&buffer, &lineLength, columns, numColumns);
      if (rc == OK) {
        printRow(columns, numColumns);
      } else if (rc != AT_END_OF_FILE) {
        fprintf(stderr, "Error %i\n", rc);
        return rc;
      }
    }
    delete[] buffer;
    delete[] columns;
  }
  return 0;
}
