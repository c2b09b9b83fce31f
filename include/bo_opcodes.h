/* bo_opcodes.h -- tape opcodes; generated from optas_b200/sym.py (tests/test_abi.py checks they agree). */
#ifndef BO_OPCODES_H
#define BO_OPCODES_H
#define BO_OP_SYM 0
#define BO_OP_CONST 1
#define BO_OP_INPUT 2
#define BO_OP_OUTPUT 3
#define BO_OP_ADD 10
#define BO_OP_SUB 11
#define BO_OP_MUL 12
#define BO_OP_DIV 13
#define BO_OP_ATAN2 14
#define BO_OP_FMIN 15
#define BO_OP_FMAX 16
#define BO_OP_POW 17
#define BO_OP_LT 18
#define BO_OP_LE 19
#define BO_OP_EQ 20
#define BO_OP_NE 21
#define BO_OP_AND 22
#define BO_OP_OR 23
#define BO_OP_NEG 40
#define BO_OP_SQ 41
#define BO_OP_SQRT 42
#define BO_OP_SIN 43
#define BO_OP_COS 44
#define BO_OP_TAN 45
#define BO_OP_ASIN 46
#define BO_OP_ACOS 47
#define BO_OP_ATAN 48
#define BO_OP_FABS 49
#define BO_OP_EXP 50
#define BO_OP_LOG 51
#define BO_OP_NOT 52
#define BO_OP_SIGN 53
#define BO_OP_FLOOR 54
#define BO_OP_CEIL 55
#define BO_OP_TANH 56
#define BO_OP_SINH 57
#define BO_OP_COSH 58
#define BO_OP_IF_ELSE 70
#endif
