/*
 * oracle/match_oracle.cpp — CPU ORACLE for the ORB matcher arithmetic (test infrastructure, NOT
 * product code).  Restates orb_slam3/src/ORBmatcher.cc of snt-arg/visual_sgraphs on flattened arrays.
 * See oracle.h for the parity status.
 */
#include "oracle.h"

#include <cstring>

extern "C" {

// ORBmatcher::DescriptorDistance — ORBmatcher.cc:2047-2063: 8 x int32 XOR + bit-parallel popcount.
int orc_descriptor_distance(const uint8_t *a, const uint8_t *b) {
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t wa, wb;
        std::memcpy(&wa, a + 4 * i, 4);
        std::memcpy(&wb, b + 4 * i, 4);
        uint32_t v = wa ^ wb;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0x0F0F0F0Fu) * 0x01010101u) >> 24);
    }
    return dist;
}

}  // extern "C"
