// buffered_output_test -- TEST INFRASTRUCTURE: the decoder's file sink (BufferedOutput of zdw_b200/host/zdw/
// UnconvertFromZDW.h) on the CPU: blocks handed over with writeLater() and small direct write()s reach the file in
// call order, a block is on its way out before its buffer is re-used two calls later (the two-context rotation of
// parseNextBlock), the destructor drains.  Exit code 0 = all good.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "zdw/UnconvertFromZDW.h"

using adobe::zdw::BufferedOutput;

int main() {
  int bad = 0;
  for (int round = 0; round < 3; ++round) {
    FILE* f = tmpfile();
    std::string want;
    {
      BufferedOutput sink(f);
      // two buffers used in turn, like the rows of the two GPU contexts
      std::vector<char> slab[2];
      for (int b = 0; b < 40; ++b) {
        if (b % 5 == 0) {  // the --non-empty-column-header line goes out directly
          const std::string head = "header " + std::to_string(b) + "\n";
          if (!sink.write(head.data(), head.size())) ++bad;
          want += head;
        }
        std::vector<char>& s = slab[b & 1];
        // re-using this slab is what the next decode call on the same context does: its previous content must be out
        s.assign(1 + (size_t)(b * 104729) % (round ? 3000000 : 5000), (char)('a' + b % 26));
        s[0] = '<';
        want.append(s.data(), s.size());
        sink.writeLater(s.data(), s.size());
      }
      sink.writeLater(NULL, 0);  // nothing to write: still waits for the block in flight
      if (round == 2 && !sink.waitIdle()) ++bad;
    }  // destructor drains
    fflush(f);
    rewind(f);
    std::string got(want.size() + 16, 0);
    const size_t n = fread(&got[0], 1, got.size(), f);
    got.resize(n);
    if (got != want) {
      printf("MISMATCH round %d: %zu bytes, %zu expected\n", round, got.size(), want.size());
      ++bad;
    }
    fclose(f);
  }
  // a sink that cannot write reports it
  {
    FILE* ro = fopen("/dev/full", "w");
    if (ro) {
      setvbuf(ro, NULL, _IONBF, 0);
      BufferedOutput sink(ro);
      std::vector<char> s(100000, 'x');
      sink.writeLater(s.data(), s.size());
      if (sink.waitIdle()) {
        printf("MISMATCH: a short write went unnoticed\n");
        ++bad;
      }
    }
    if (ro) fclose(ro);
  }
  printf("%d mismatches\n", bad);
  return bad ? 1 : 0;
}
