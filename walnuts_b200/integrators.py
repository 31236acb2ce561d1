"""Macro-step integrator handles, mirroring reference WALNUTSpy/adaptiveIntegrators.py.

The reference passes Python functions `integrator(q, v, g, Ham0, h, xi, lpFun, delta, auxPar)`
(adaptiveIntegrators.py:49,65,361); here the same names are registry handles selecting the CUDA
implementation inside the persistent kernel (walnuts_b200/csrc/wn_walnutspy.cuh).
"""


class _Integrator:
    def __init__(self, name, kind, ref):
        self.name, self.kind, self.ref = name, kind, ref

    def __repr__(self):
        return f"<walnuts_b200 integrator {self.name} ({self.ref})>"

    def __call__(self, *a, **k):
        raise TypeError(f"{self.name} is a handle for the CUDA integrator; it runs inside "
                        "walnuts_b200.WALNUTS(...) and cannot be called on host arrays")


fixedLeapFrog = _Integrator("fixedLeapFrog", 0, "adaptiveIntegrators.py:49-59")
adaptLeapFrogD = _Integrator("adaptLeapFrogD", 1, "adaptiveIntegrators.py:65-137")
adaptLeapFrogR2P = _Integrator("adaptLeapFrogR2P", 2, "adaptiveIntegrators.py:361-475")
adaptYoshidaD = _Integrator("adaptYoshidaD", 3, "adaptiveIntegrators.py:142-240")
adaptLeapFrogFlowD = _Integrator("adaptLeapFrogFlowD", 4, "adaptiveIntegrators.py:246-356")
adaptImplicitMidpointD = _Integrator("adaptImplicitMidpointD", 5, "adaptiveIntegrators.py:478-641")
adaptRescaledLeapFrogD = _Integrator("adaptRescaledLeapFrogD", 6, "adaptiveIntegrators.py:660-762")
KINDS = {"fixed": 0, "D": 1, "R2P": 2, "Yoshida": 3, "Flow": 4, "Midpoint": 5, "Rescaled": 6}


class integratorAuxPar:
    """adaptiveIntegrators.integratorAuxPar (adaptiveIntegrators.py:36-44).  maxFPiter / FPtol steer
    adaptImplicitMidpointD, rescaledGradThresh steers adaptRescaledLeapFrogD; FPNewton=True (Newton iterations on
    the target's Hessian, adaptiveIntegrators.py:503-506) has no CUDA implementation and is refused."""

    def __init__(self, minC=0, maxC=10, R2Pprob0=2.0 / 3.0, maxFPiter=30, FPtol=1.0e-8, FPNewton=False,
                 rescaledGradThresh=5.0):
        self.minC = minC
        self.maxC = maxC
        self.R2Pprob0 = R2Pprob0
        self.maxFPiter = maxFPiter
        self.FPtol = FPtol
        self.FPNewton = FPNewton
        self.rescaledGradThresh = rescaledGradThresh
