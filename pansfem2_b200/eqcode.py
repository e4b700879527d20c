"""PF2_EQ_CODE of include/pansfem2_b200.h on the Python side: build and describe element selections
<Equation, ShapeFunction, Integration> (host-side bookkeeping only)."""
from __future__ import annotations

PHYS_PLANESTRAIN, PHYS_SOLID, PHYS_HEAT, PHYS_PLANESTRESS, PHYS_PLANESTRAIN_SRI, PHYS_MASS, PHYS_PLANESTRAIN_BBAR, PHYS_MASS2, PHYS_PLANESTRAIN_WT, PHYS_ADVDIFF, PHYS_PLANE_D, PHYS_PLANE_D_BBAR, PHYS_PLANE_D_WT = range(13)
# PF2_ADV_* routine mask of PHYS_ADVDIFF (carried in the quad2 field): Advection.h:19, :135, :47, :91, :161, :188
ADV_ADVECTION, ADV_DIFFUSION, ADV_SUPG, ADV_SHOCK, ADV_MASS, ADV_MASS_SUPG = 1, 2, 4, 8, 16, 32
ADV_NAME = {1: "Advection", 2: "Diffusion", 4: "AdvectionSUPG", 8: "AdvectionShockCapturing", 16: "Mass", 32: "MassSUPG"}
SHAPE_DEFAULT, SHAPE_T3, SHAPE_T6, SHAPE_Q4, SHAPE_Q8, SHAPE_TET4, SHAPE_HEX8, SHAPE_HEX20 = range(8)
QUAD_DEFAULT, QUAD_G1TRI, QUAD_G3TRI, QUAD_G1SQ, QUAD_G4SQ, QUAD_G9SQ, QUAD_G1TET, QUAD_G8CUBE, QUAD_G27CUBE = range(9)

SHAPE_NPE = {SHAPE_T3: 3, SHAPE_T6: 6, SHAPE_Q4: 4, SHAPE_Q8: 8, SHAPE_TET4: 4, SHAPE_HEX8: 8, SHAPE_HEX20: 20}
SHAPE_NAME = {SHAPE_T3: "T3", SHAPE_T6: "T6", SHAPE_Q4: "Q4", SHAPE_Q8: "Q8", SHAPE_TET4: "Tet4", SHAPE_HEX8: "Hex8", SHAPE_HEX20: "Hex20"}
QUAD_NAME = {QUAD_G1TRI: "Gauss1Triangle", QUAD_G3TRI: "Gauss3Triangle", QUAD_G1SQ: "Gauss1Square", QUAD_G4SQ: "Gauss4Square",
             QUAD_G9SQ: "Gauss9Square", QUAD_G1TET: "Gauss1Tetrahedron", QUAD_G8CUBE: "Gauss8Cubic", QUAD_G27CUBE: "Gauss27Cubic"}
PHYS_NAME = {PHYS_PLANESTRAIN: "PlaneStrain", PHYS_SOLID: "Solid", PHYS_HEAT: "HeatTransfer", PHYS_PLANESTRESS: "PlaneStress",
             PHYS_PLANESTRAIN_SRI: "PlaneStrainSRI", PHYS_MASS: "ConsistentMass", PHYS_PLANESTRAIN_BBAR: "PlaneStrainBbar", PHYS_MASS2: "ConsistentMass2dof", PHYS_PLANESTRAIN_WT: "PlaneStrainWilsonTaylor",
             PHYS_ADVDIFF: "AdvectionDiffusion", PHYS_PLANE_D: "PlaneStiffness", PHYS_PLANE_D_BBAR: "PlaneStiffnessBbar",
             PHYS_PLANE_D_WT: "PlaneStiffnessWilsonTaylor"}
# rules of each reference domain (triangle, square, tetrahedron, cube)
SHAPE_RULES = {SHAPE_T3: (QUAD_G1TRI, QUAD_G3TRI), SHAPE_T6: (QUAD_G1TRI, QUAD_G3TRI),
               SHAPE_Q4: (QUAD_G1SQ, QUAD_G4SQ, QUAD_G9SQ), SHAPE_Q8: (QUAD_G1SQ, QUAD_G4SQ, QUAD_G9SQ),
               SHAPE_TET4: (QUAD_G1TET,), SHAPE_HEX8: (QUAD_G8CUBE, QUAD_G27CUBE), SHAPE_HEX20: (QUAD_G8CUBE, QUAD_G27CUBE)}
DEFAULT_RULE = {SHAPE_T3: QUAD_G1TRI, SHAPE_T6: QUAD_G1TRI, SHAPE_Q4: QUAD_G4SQ, SHAPE_Q8: QUAD_G4SQ, SHAPE_TET4: QUAD_G1TET,
                SHAPE_HEX8: QUAD_G8CUBE, SHAPE_HEX20: QUAD_G8CUBE}


def eq_code(phys, shape=0, quad=0, quad2=0) -> int:
    return phys | (shape << 8) | (quad << 16) | (quad2 << 24)


def fields(eq):
    """(phys, shape, quad, quad2) with the defaults of the physics filled in (mirrors decode_eq, csrc/assemble_generic.cu)."""
    phys, shape, quad, quad2 = eq & 0xff, (eq >> 8) & 0xff, (eq >> 16) & 0xff, (eq >> 24) & 0xff
    solid = phys == PHYS_SOLID
    if shape == 0:
        shape = SHAPE_HEX8 if solid else SHAPE_Q4
    if quad == 0:
        quad = DEFAULT_RULE[shape]
    if phys in (PHYS_PLANESTRAIN_SRI, PHYS_PLANESTRAIN_BBAR, PHYS_PLANE_D_BBAR) and quad2 == 0:
        quad2 = QUAD_G1TRI if shape in (SHAPE_T3, SHAPE_T6) else QUAD_G1SQ
    return phys, shape, quad, quad2


def ndof(eq) -> int:
    phys = eq & 0xff
    return 3 if phys == PHYS_SOLID else (1 if phys in (PHYS_HEAT, PHYS_MASS, PHYS_ADVDIFF) else 2)


def dim(eq) -> int:
    return 3 if (eq & 0xff) == PHYS_SOLID else 2


def npe(eq) -> int:
    return SHAPE_NPE[fields(eq)[1]]


def describe(eq) -> str:
    phys, shape, quad, quad2 = fields(eq)
    if phys == PHYS_ADVDIFF:
        return "+".join(n for b, n in ADV_NAME.items() if quad2 & b) + f"<{SHAPE_NAME[shape]},{QUAD_NAME[quad]}>"
    s = f"{PHYS_NAME[phys]}<{SHAPE_NAME[shape]},{QUAD_NAME[quad]}"
    return s + (f",{QUAD_NAME[quad2]}>" if quad2 else ">")
