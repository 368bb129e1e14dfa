"""B200-native drop-ins for the reference's ``src/models`` modules (same module and class names, so
``config.model.module = "gnn_fluid_dynamics_b200.models.Fvgn"``, ``config.model.name = "FvgnA"``)."""
from .Fvgn import FvgnA, FvgnB, FvgnC, FvgnD, FvgnE, FvgnF, FvgnH, FvgnI, FvgnJ, FvgnK  # noqa: F401
from .Mgn import MgnA, MgnB, MgnC  # noqa: F401
from .StreamFunc import StreamFuncA, StreamFuncB, StreamFuncC, StreamFuncD  # noqa: F401
from .Flux import FluxA, FluxB, FluxC, FluxD  # noqa: F401
from .Conservative import (ConservativeA, ConservativeB, ConservativeD, ConservativeE, ConservativeF, ConservativeG,  # noqa: F401
                           ConservativeH, ConservativeI, ConservativeJ, ConservativeK)
from .VertPot import VertPotA, VertPotB, VertPotC, VertPotD, VertPotE, VertPotF, VertPotG  # noqa: F401

# VertPotD / VertPotF are importable (constructor, state_dict, classmethods) but their forward raises: it does in the
# reference too (missing function), so they are not in MODEL_CLASSES (the classes with a reference behaviour to match)
UNRUNNABLE_CLASSES = {"VertPotD": VertPotD, "VertPotF": VertPotF}
MODEL_CLASSES = {"FvgnA": FvgnA, "FvgnF": FvgnF, "MgnA": MgnA, "FluxA": FluxA, "ConservativeA": ConservativeA,
                 "VertPotA": VertPotA, "ConservativeE": ConservativeE, "ConservativeF": ConservativeF, "ConservativeD": ConservativeD,
                 "ConservativeG": ConservativeG, "ConservativeI": ConservativeI, "ConservativeH": ConservativeH,
                 "ConservativeK": ConservativeK, "MgnB": MgnB, "MgnC": MgnC, "StreamFuncA": StreamFuncA,
                 "ConservativeB": ConservativeB, "ConservativeJ": ConservativeJ, "VertPotB": VertPotB, "VertPotC": VertPotC, "VertPotE": VertPotE, "VertPotG": VertPotG, "FvgnB": FvgnB, "FvgnC": FvgnC, "FvgnD": FvgnD, "FvgnE": FvgnE, "FvgnH": FvgnH, "FvgnI": FvgnI, "FvgnJ": FvgnJ,
                 "FvgnK": FvgnK, "FluxB": FluxB, "FluxC": FluxC, "FluxD": FluxD, "StreamFuncB": StreamFuncB, "StreamFuncC": StreamFuncC, "StreamFuncD": StreamFuncD}
