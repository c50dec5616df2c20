#include "zdw/status_output.h"

#include <cstdarg>
#include <cstdio>

namespace adobe {
namespace zdw {

namespace {
void emit(FILE* to, const char* format, va_list ap) {
  vfprintf(to, format, ap);
  fflush(to);
}
}  // namespace

void defaultStatusOutputCallback(const StatusOutputLevel level, const char* format, ...) {
  va_list ap;
  va_start(ap, format);
  emit(level == ERROR ? stderr : stdout, format, ap);
  va_end(ap);
}

void stdErrStatusOutputCallback(const StatusOutputLevel, const char* format, ...) {
  va_list ap;
  va_start(ap, format);
  emit(stderr, format, ap);
  va_end(ap);
}

}  // namespace zdw
}  // namespace adobe
