#pragma once
#include <emmintrin.h>
