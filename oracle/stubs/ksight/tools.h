/* Stand-in for Falcon's un-vendored ksight profiling header, so that the reference's
 * pairhmm/interface/PairHMMHostInterface.cpp compiles from where it lies (oracle/Makefile, _ref/libhostif_ref.so).
 * The reference only uses the two timer macros; they expand to nothing.  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_STUB_KSIGHT_TOOLS_H
#define ORACLE_STUB_KSIGHT_TOOLS_H
#define PLACE_TIMER
#define PLACE_TIMER1(x)
#endif
