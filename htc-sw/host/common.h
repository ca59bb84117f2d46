// common.h -- constants and result types of the Smith-Waterman host interface, with the names and values of the
// reference's header (/root/reference/htc-sw/host/common.h:9-51) so that its callers compile against this one.
#ifndef COMMON_H
#define COMMON_H

#define MAX_SEQ_LENGTH 1536
#define MAX_BATCH_SIZE 260
#define OVERHANG_STRATEGY_SOFTCLIP 0
#define OVERHANG_STRATEGY_INDEL 1
#define OVERHANG_STRATEGY_LEADING_INDEL 2
#define OVERHANG_STRATEGY_IGNORE 3
#define W_MATCH 200
#define W_MISMATCH -150
#define W_OPEN -260
#define W_EXTEND -11
#define STATE_MATCH 0       // CigarOperator.M
#define STATE_INSERTION 1   // CigarOperator.I
#define STATE_DELETION 2    // CigarOperator.D
#define STATE_CLIP 4        // CigarOperator.S

struct CigarElement {
    int length;
    int state;
};

struct Cigar {
    struct CigarElement cigarElements[MAX_SEQ_LENGTH];
    int CigarElementNum;
};

#endif
