// status_output.h -- printf-style status sink shared by the encoder and decoder host classes.
// Same contract as the reference's zdw/status_output.h:17-33: INFO lines go to stdout and ERROR lines to stderr by
// default; the "-" (TSV to stdout) mode of unconvertDWfile swaps in the all-to-stderr sink.
#ifndef ZDWB_HOST_STATUS_OUTPUT_H
#define ZDWB_HOST_STATUS_OUTPUT_H

namespace adobe {
namespace zdw {

enum StatusOutputLevel { INFO, ERROR };

typedef void (*StatusOutputCallback)(const StatusOutputLevel, const char*, ...);

void defaultStatusOutputCallback(const StatusOutputLevel level, const char* format, ...);
void stdErrStatusOutputCallback(const StatusOutputLevel level, const char* format, ...);

}  // namespace zdw
}  // namespace adobe
#endif
