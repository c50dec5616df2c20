// includes.h -- integer aliases and column type ids used across the host layer.
// The ids are the on-disk values of the file header's type vector (reference: zdw_column_type_constants.h:17-38);
// they are spelled as an enum here and mirrored by ZDWB_* in include/zdw_b200.h.
#ifndef ZDWB_HOST_INCLUDES_H
#define ZDWB_HOST_INCLUDES_H

#include <stdint.h>

namespace adobe {
namespace zdw {

typedef uint32_t ULONG;
typedef uint16_t USHORT;
typedef uint8_t UCHAR;
typedef uint64_t ULONGLONG;
typedef int64_t SLONGLONG;

enum ColumnTypeId {
  ZT_VARCHAR = 0, ZT_TEXT = 1, ZT_DATETIME = 2, ZT_CHAR_2 = 3, ZT_VISID_LOW = 4, ZT_VISID_HIGH = 5, ZT_CHAR = 6,
  ZT_TINY = 7, ZT_SHORT = 8, ZT_LONG = 9, ZT_LONGLONG = 10, ZT_DECIMAL = 11, ZT_TINY_SIGNED = 12,
  ZT_SHORT_SIGNED = 13, ZT_LONG_SIGNED = 14, ZT_LONGLONG_SIGNED = 15, ZT_TINYTEXT = 16, ZT_MEDIUMTEXT = 17,
  ZT_LONGTEXT = 18,
  // in-memory only, never stored (reference :36-38)
  ZT_VIRTUAL_EXPORT_FILE_BASENAME = 64, ZT_VIRTUAL_EXPORT_ROW = 65
};

}  // namespace zdw
}  // namespace adobe
#endif
