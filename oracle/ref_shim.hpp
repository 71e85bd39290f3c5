// ref_shim.hpp -- force-included (-include) into every translation unit of oracle/build_ref.py: the two clang builtins
// g++ 13 lacks (16-bit count-leading / trailing zeros, used by floor's bit-math helpers, not on the image path).
// TEST INFRASTRUCTURE ONLY.
#pragma once
constexpr int floor_ref_clzs(unsigned short v) { return v == 0 ? 16 : __builtin_clz((unsigned int)v) - 16; }
constexpr int floor_ref_ctzs(unsigned short v) { return v == 0 ? 16 : __builtin_ctz((unsigned int)v); }
