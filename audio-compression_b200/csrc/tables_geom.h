// tables_geom.h — block geometry of tables.cu, shared with the CPU emulation in tests/csrc/host_harness.cpp.
#pragma once

constexpr int kChainT = 17;                          // chains per thread (17 = 1 mod 4: the four 8-lane groups of a warp hit disjoint banks)
constexpr int kChainThreads = 256;
constexpr int kChainP = kChainT * kChainThreads;     // 4352 chain starts per block
constexpr int kChainOut = kChainP - 8;               // leaf starts a block serves: [Q0, Q0 + 4344); a leaf needs chains p..p+7
constexpr int kChainX = kChainP + 120;               // samples staged per block (the last chain reads 120 past its start)
constexpr int kTabThreads = 128;
constexpr int kTabJ = 512;                           // domains per block pass of tables_from_halves_kernel
