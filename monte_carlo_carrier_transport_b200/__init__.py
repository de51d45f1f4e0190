"""B200-native ensemble Monte Carlo carrier transport (hot path of
edsolibet/Monte_Carlo_Carrier_Transport behind the reference's module surface).

    init_, scatter_, main_, poisson_   -- mirrors of the reference's modules
    engine.Ensemble                    -- batched device-resident ensemble (the fast path)
    abi                                -- ctypes binding of the C-ABI in include/emc.h

The compute path is hand-written CUDA for sm_100a in csrc/, reached through a C-ABI
shared library.  There is no CPU fallback: any call that needs the GPU raises if the
library or a CUDA device is missing.
"""
__version__ = "0.1.0"
