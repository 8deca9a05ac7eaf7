// Shim: the reference includes simde only as a portable alias of the x86 intrinsics.
#pragma once
#include <xmmintrin.h>
