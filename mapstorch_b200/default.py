"""Parameter / element dictionaries of the fitting API (boundary data).

Same names and values as the reference's ``mapstorch/default.py:51-373`` so that user
code which imports them keeps working; they are the inputs of the line-table compiler
(:mod:`mapstorch_b200.tables`) that feeds the CUDA kernels.  Values are stated once,
in table form, and expanded at import.
"""
from mapstorch_b200.constant import MapsElements, define_constants

# name -> (default value, tunable?, learning rate or None)      default.py:51-116
_PARAM_TABLE = """
GE_ESCAPE             0.0                       -  -
COMPTON_F_TAIL        0.783374011516571         T  1e-4
COMPTON_ANGLE         109.85600280761719        T  1
KB_F_TAIL_QUADRATIC   0.0                       -  -
COHERENT_SCT_ENERGY   12.270400047302246        T  1e-2
F_TAIL_OFFSET         0.03593109920620918       T  1e-4
F_STEP_LINEAR         0.0                       -  -
GAMMA_QUADRATIC       0.0                       -  -
FWHM_OFFSET           0.08414760231971741       T  1e-4
ENERGY_QUADRATIC      -2.2838600344243787e-08   T  1e-7
COMPTON_FWHM_CORR     2.7478699684143066        T  1e-2
F_STEP_QUADRATIC      0.0                       -  -
COMPTON_F_STEP        0.0                       -  -
SI_ESCAPE             0.0                       -  -
FWHM_FANOPRIME        0.00017760999617166817    T  1e-8
COMPTON_HI_GAMMA      3.0                       -  -
ENERGY_OFFSET         -0.0020250400993973017    T  1e-3
KB_F_TAIL_OFFSET      0.16684700548648834       T  1e-4
ENERGY_SLOPE          0.011979999952018261      T  1e-6
F_TAIL_QUADRATIC      0.0                       -  -
COMPTON_GAMMA         3.0                       -  -
F_STEP_OFFSET         0.0                       -  -
COMPTON_HI_F_TAIL     0.08728790283203125       T  1e-6
F_TAIL_LINEAR         1.6940699334764354e-21    -  -
GAMMA_OFFSET          2.2101199626922607        -  -
KB_F_TAIL_LINEAR      0.0                       -  -
MAX_ENERGY_TO_FIT     14.800000190734863        -  -
GAMMA_LINEAR          0.0                       -  -
MIN_ENERGY_TO_FIT     1.0                       -  -
SNIP_WIDTH            0.5                       -  -
"""

# the order in which the reference lists its tunable parameters (default.py:51-64)
_TUNABLE_ORDER = ("COHERENT_SCT_ENERGY ENERGY_OFFSET ENERGY_SLOPE ENERGY_QUADRATIC COMPTON_ANGLE "
                  "COMPTON_FWHM_CORR COMPTON_HI_F_TAIL COMPTON_F_TAIL FWHM_FANOPRIME FWHM_OFFSET "
                  "F_TAIL_OFFSET KB_F_TAIL_OFFSET").split()

default_param_vals = {}
default_learning_rates = {"default_lr": 1e-2}
_tunable = set()
for _row in _PARAM_TABLE.strip().splitlines():
    _name, _val, _flag, _lr = _row.split()
    default_param_vals[_name] = float(_val)
    if _flag == "T":
        _tunable.add(_name)
        default_learning_rates[_name] = float(_lr)
assert _tunable == set(_TUNABLE_ORDER)
default_fitting_params = list(_TUNABLE_ORDER)
# keep the reference's key order for the learning-rate dict (default.py:102-116)
default_learning_rates = {k: default_learning_rates[k] for k in (
    "default_lr COHERENT_SCT_ENERGY ENERGY_OFFSET ENERGY_SLOPE ENERGY_QUADRATIC COMPTON_ANGLE "
    "COMPTON_FWHM_CORR COMPTON_HI_F_TAIL COMPTON_F_TAIL FWHM_FANOPRIME FWHM_OFFSET F_TAIL_OFFSET "
    "KB_F_TAIL_OFFSET").split()}

# MAPS override-file tag -> parameter name (default.py:118-131)
param_name_map = dict(pair.split("=") for pair in (
    "CAL_OFFSET_[E_OFFSET]=ENERGY_OFFSET CAL_SLOPE_[E_LINEAR]=ENERGY_SLOPE "
    "CAL_QUAD_[E_QUADRATIC]=ENERGY_QUADRATIC STEP_OFFSET=F_STEP_OFFSET STEP_LINEAR=F_STEP_LINEAR "
    "STEP_QUADRATIC=F_STEP_QUADRATIC COMPTON_OFFSET=COMPTON_F_STEP COMPTON_LINEAR=COMPTON_F_LINEAR "
    "COMPTON_QUADRATIC=COMPTON_F_QUADRATIC COMPTON_STEP=COMPTON_F_STEP SI_ESCAPE_ENABLE=SI_ESCAPE "
    "GE_ESCAPE_ENABLE=GE_ESCAPE").split())

# element -> which line families are fitted by default (default.py:134-227, 305-367)
supported_elements_mapping = {sym: list(fam) for sym, fam in (item.split(":") for item in (
    "Ag:KL Al:K Ar:K As:KL Au:LM Ba:L Bi:L Br:KL Ca:K Cd:KL Ce:L Cl:K Co:K Cr:K Cs:L Cu:K Dy:L "
    "Er:L Eu:L Fe:K Ga:K Gd:L Ge:K Hf:KLM Hg:LM Ho:L I:KL K:K Kr:L La:L Lu:L Mg:K Mn:K Mo:KL "
    "Nd:L Ni:K Os:LM P:K Pa:L Pb:KL Pd:K Pm:L Pr:L Pt:LM Rb:KL S:K Sc:K Se:KL Si:K Sm:L Sn:L "
    "Sr:KL Ta:L Tb:L Th:LM Ti:K Tl:L Tm:L U:LM V:K W:LM Xe:L Y:KL Yb:L Zn:KL Zr:KL").split())}

default_K_lines = [s for s, fam in supported_elements_mapping.items() if "K" in fam]
default_L_lines = [s + "_L" for s, fam in supported_elements_mapping.items() if "L" in fam]
default_M_lines = [s + "_M" for s, fam in supported_elements_mapping.items() if "M" in fam]
default_pileups = ["Si_Si", "Si_Cl", "Cl_Cl"]
default_auxiliary_comps = ["COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE"]
default_fitting_elems = (default_K_lines + default_pileups + default_auxiliary_comps
                         + default_L_lines + default_M_lines)

unsupported_elements = (
    "Ac Am At B Be Bh Bk C Cf Cm Cn Db Ds Es F Fl Fm Fr H He Hs In Ir Li Lr Lv Mc Md Mt N Na Nb "
    "Ne Nh No Np O Og Po Pu Ra Re Rf Rg Rh Rn Ru Sb Sg Tc Te Ts").split()

# element information (per Z) and the per-element emission-line tables        default.py:370-373
default_elem_info = MapsElements().get_element_info()
default_energy_consts = define_constants(default_elem_info, default_pileups)
