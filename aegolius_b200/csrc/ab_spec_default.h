// ab_spec_default.h — per-op compile switches of the interpreter. Every op is compiled in unless a specialised build
// (aegolius_b200/build.py: build_specialized) passes -DAB_SPEC_<OP>=0 for the ops a program does not use.
#pragma once
#ifndef AB_SPEC_END
#define AB_SPEC_END 1
#endif
#ifndef AB_SPEC_SAVE_P
#define AB_SPEC_SAVE_P 1
#endif
#ifndef AB_SPEC_LOAD_P
#define AB_SPEC_LOAD_P 1
#endif
#ifndef AB_SPEC_PUSH_V
#define AB_SPEC_PUSH_V 1
#endif
#ifndef AB_SPEC_NEXT_AFFINE
#define AB_SPEC_NEXT_AFFINE 1
#endif
#ifndef AB_SPEC_NEXT_TRANSLATE
#define AB_SPEC_NEXT_TRANSLATE 1
#endif
#ifndef AB_SPEC_NEXT_LOAD
#define AB_SPEC_NEXT_LOAD 1
#endif
#ifndef AB_SPEC_AFFINE
#define AB_SPEC_AFFINE 1
#endif
#ifndef AB_SPEC_TRANSLATE
#define AB_SPEC_TRANSLATE 1
#endif
#ifndef AB_SPEC_SCALE_P
#define AB_SPEC_SCALE_P 1
#endif
#ifndef AB_SPEC_ELONGATE
#define AB_SPEC_ELONGATE 1
#endif
#ifndef AB_SPEC_TWIST
#define AB_SPEC_TWIST 1
#endif
#ifndef AB_SPEC_BEND
#define AB_SPEC_BEND 1
#endif
#ifndef AB_SPEC_ABSX_SUB
#define AB_SPEC_ABSX_SUB 1
#endif
#ifndef AB_SPEC_SYMMETRY
#define AB_SPEC_SYMMETRY 1
#endif
#ifndef AB_SPEC_ROTSYM
#define AB_SPEC_ROTSYM 1
#endif
#ifndef AB_SPEC_REVOLVE
#define AB_SPEC_REVOLVE 1
#endif
#ifndef AB_SPEC_AXIS_REVOLVE
#define AB_SPEC_AXIS_REVOLVE 1
#endif
#ifndef AB_SPEC_REP_INF
#define AB_SPEC_REP_INF 1
#endif
#ifndef AB_SPEC_REP_FIN
#define AB_SPEC_REP_FIN 1
#endif
#ifndef AB_SPEC_LIN_INST
#define AB_SPEC_LIN_INST 1
#endif
#ifndef AB_SPEC_CURVE_INST
#define AB_SPEC_CURVE_INST 1
#endif
#ifndef AB_SPEC_ZERO_Z
#define AB_SPEC_ZERO_Z 1
#endif
#ifndef AB_SPEC_ROUND
#define AB_SPEC_ROUND 1
#endif
#ifndef AB_SPEC_ABS
#define AB_SPEC_ABS 1
#endif
#ifndef AB_SPEC_NEG
#define AB_SPEC_NEG 1
#endif
#ifndef AB_SPEC_SIGN
#define AB_SPEC_SIGN 1
#endif
#ifndef AB_SPEC_ONION
#define AB_SPEC_ONION 1
#endif
#ifndef AB_SPEC_CONCENTRIC
#define AB_SPEC_CONCENTRIC 1
#endif
#ifndef AB_SPEC_SCALE_V
#define AB_SPEC_SCALE_V 1
#endif
#ifndef AB_SPEC_EXTRUDE_BEGIN
#define AB_SPEC_EXTRUDE_BEGIN 1
#endif
#ifndef AB_SPEC_EXTRUDE_END
#define AB_SPEC_EXTRUDE_END 1
#endif
#ifndef AB_SPEC_PP_SIGMOID
#define AB_SPEC_PP_SIGMOID 1
#endif
#ifndef AB_SPEC_PP_POS_SIGMOID
#define AB_SPEC_PP_POS_SIGMOID 1
#endif
#ifndef AB_SPEC_PP_CAPPED_EXP
#define AB_SPEC_PP_CAPPED_EXP 1
#endif
#ifndef AB_SPEC_PP_HARD_BIN
#define AB_SPEC_PP_HARD_BIN 1
#endif
#ifndef AB_SPEC_PP_LINEAR
#define AB_SPEC_PP_LINEAR 1
#endif
#ifndef AB_SPEC_PP_RELU
#define AB_SPEC_PP_RELU 1
#endif
#ifndef AB_SPEC_PP_SMOOTH_RELU
#define AB_SPEC_PP_SMOOTH_RELU 1
#endif
#ifndef AB_SPEC_PP_SLOWSTART
#define AB_SPEC_PP_SLOWSTART 1
#endif
#ifndef AB_SPEC_PP_GAUSS_BOUNDARY
#define AB_SPEC_PP_GAUSS_BOUNDARY 1
#endif
#ifndef AB_SPEC_PP_GAUSS_FALLOFF
#define AB_SPEC_PP_GAUSS_FALLOFF 1
#endif
#ifndef AB_SPEC_C_UNION
#define AB_SPEC_C_UNION 1
#endif
#ifndef AB_SPEC_C_INTERSECT
#define AB_SPEC_C_INTERSECT 1
#endif
#ifndef AB_SPEC_C_SUBTRACT
#define AB_SPEC_C_SUBTRACT 1
#endif
#ifndef AB_SPEC_C_SUM
#define AB_SPEC_C_SUM 1
#endif
#ifndef AB_SPEC_C_DIFF
#define AB_SPEC_C_DIFF 1
#endif
#ifndef AB_SPEC_C_SMIN2
#define AB_SPEC_C_SMIN2 1
#endif
#ifndef AB_SPEC_C_SMIN3
#define AB_SPEC_C_SMIN3 1
#endif
#ifndef AB_SPEC_C_SMAX3
#define AB_SPEC_C_SMAX3 1
#endif
#ifndef AB_SPEC_C_SSUB3
#define AB_SPEC_C_SSUB3 1
#endif
#ifndef AB_SPEC_C_BOLTZ_INT
#define AB_SPEC_C_BOLTZ_INT 1
#endif
#ifndef AB_SPEC_C_BOLTZ_SUB
#define AB_SPEC_C_BOLTZ_SUB 1
#endif
#ifndef AB_SPEC_P_SPHERE
#define AB_SPEC_P_SPHERE 1
#endif
#ifndef AB_SPEC_P_CYLINDER
#define AB_SPEC_P_CYLINDER 1
#endif
#ifndef AB_SPEC_P_BOX
#define AB_SPEC_P_BOX 1
#endif
#ifndef AB_SPEC_P_TORUS
#define AB_SPEC_P_TORUS 1
#endif
#ifndef AB_SPEC_P_CHAINLINK
#define AB_SPEC_P_CHAINLINK 1
#endif
#ifndef AB_SPEC_P_BRAID
#define AB_SPEC_P_BRAID 1
#endif
#ifndef AB_SPEC_P_ARC3D
#define AB_SPEC_P_ARC3D 1
#endif
#ifndef AB_SPEC_P_PLANE
#define AB_SPEC_P_PLANE 1
#endif
#ifndef AB_SPEC_P_UPLANE
#define AB_SPEC_P_UPLANE 1
#endif
#ifndef AB_SPEC_P_SEGMENT
#define AB_SPEC_P_SEGMENT 1
#endif
#ifndef AB_SPEC_P_CONE
#define AB_SPEC_P_CONE 1
#endif
#ifndef AB_SPEC_P_OINF_CONE
#define AB_SPEC_P_OINF_CONE 1
#endif
#ifndef AB_SPEC_P_INF_CONE
#define AB_SPEC_P_INF_CONE 1
#endif
#ifndef AB_SPEC_P_SOLID_ANGLE
#define AB_SPEC_P_SOLID_ANGLE 1
#endif
#ifndef AB_SPEC_P_TRIANGLE3D
#define AB_SPEC_P_TRIANGLE3D 1
#endif
#ifndef AB_SPEC_P_QUAD3D
#define AB_SPEC_P_QUAD3D 1
#endif
#ifndef AB_SPEC_P_SEGLINE
#define AB_SPEC_P_SEGLINE 1
#endif
#ifndef AB_SPEC_P_AXIS
#define AB_SPEC_P_AXIS 1
#endif
#ifndef AB_SPEC_P_POINT_CLOUD
#define AB_SPEC_P_POINT_CLOUD 1
#endif
#ifndef AB_SPEC_P_FIELD
#define AB_SPEC_P_FIELD 1
#endif
#ifndef AB_SPEC_P_CIRCLE
#define AB_SPEC_P_CIRCLE 1
#endif
#ifndef AB_SPEC_P_NEU_CIRCLE
#define AB_SPEC_P_NEU_CIRCLE 1
#endif
#ifndef AB_SPEC_P_BOX2D
#define AB_SPEC_P_BOX2D 1
#endif
#ifndef AB_SPEC_P_SEGMENT2D
#define AB_SPEC_P_SEGMENT2D 1
#endif
#ifndef AB_SPEC_P_RBOX2D
#define AB_SPEC_P_RBOX2D 1
#endif
#ifndef AB_SPEC_P_TRIANGLE2D
#define AB_SPEC_P_TRIANGLE2D 1
#endif
#ifndef AB_SPEC_P_ARC
#define AB_SPEC_P_ARC 1
#endif
#ifndef AB_SPEC_P_SECTOR
#define AB_SPEC_P_SECTOR 1
#endif
#ifndef AB_SPEC_P_INF_SECTOR
#define AB_SPEC_P_INF_SECTOR 1
#endif
#ifndef AB_SPEC_P_NGON
#define AB_SPEC_P_NGON 1
#endif
#ifndef AB_SPEC_P_SEGLINE2D
#define AB_SPEC_P_SEGLINE2D 1
#endif
#ifndef AB_SPEC_P_POLYGON2D
#define AB_SPEC_P_POLYGON2D 1
#endif
