// scan_batch.cu — placeholder translation unit (kernel 2 lands here).
#include "internal.h"
namespace cqs {}
