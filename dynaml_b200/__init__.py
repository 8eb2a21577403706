"""dynaml_b200 -- a B200-native (sm_100a) drop-in for DynaML's Gaussian-process hot path.

Layout: csrc/ (CUDA kernels + the C ABI of include/dgp_b200.h), _capi (ctypes binding),
kernels / models / optimization (host-side mirror of the dynaml-core classes on the path).
Importing the package never touches the GPU; the first Context() does, and raises if the CUDA
library or an sm_100 device is missing (there is no CPU fallback).
"""
from . import _capi
from ._capi import Context, DgpError, NotPositiveDefinite, default_context
from .kernels import (AdditiveCovariance, CauchyKernel, CompositeCovariance, DecomposableCovariance, DiracKernel,
                      FBMKernel,
                      GenericMaternKernel, PolynomialKernel,
                      LaplacianKernel, LocalScalarKernel, MahalanobisKernel, MultiplicativeCovariance, PeriodicKernel,
                      RationalQuadraticKernel, RBFKernel, ScaledKernel, SEKernel, SVMKernelMatrix, TStudentKernel,
                      KroneckerProductKernel, TensorCombinationKernel)
from .kernels import SVMKernel
from .algebra import LowerTriPartitionedMatrix, PartitionedMatrix, PartitionedPSDMatrix, PartitionedVector
from .models import (AbstractGPRegressionModel, CudaGPRegression, GPBasisFuncRegression, GPRegression,
                     MultGaussianPRV)
from .stp import AbstractSTPRegressionModel, MultStudentsTPRV, STPRegression
from .optimization import (CoupledSimulatedAnnealing, GradBasedGlobalOptimizer, GridSearch, ModelTuner,
                           ProbGPCommMachine)
from .mcmc import AdaptiveHyperParameterMCMC, HyperParameterMCMC, HyperParameterSCAM
from .esgp import BlockedMESNRV, ESGPModel
from .mogp import KroneckerMOGPModel, MOGPRegressionModel
from .kernels import (CoRegCauchyKernel, CoRegDiracKernel, CoRegLaplaceKernel, CoRegRBFKernel,
                      CoRegTStudentKernel)

__all__ = [
    "Context", "DgpError", "NotPositiveDefinite", "default_context",
    "LocalScalarKernel", "SEKernel", "RBFKernel", "MahalanobisKernel", "GenericMaternKernel", "PeriodicKernel",
    "DiracKernel", "AdditiveCovariance", "MultiplicativeCovariance", "CompositeCovariance", "ScaledKernel",
    "SVMKernelMatrix", "SVMKernel", "PartitionedMatrix", "PartitionedPSDMatrix", "LowerTriPartitionedMatrix",
    "PartitionedVector",
    "CauchyKernel", "LaplacianKernel", "RationalQuadraticKernel", "TStudentKernel", "PolynomialKernel", "FBMKernel",
    "GPRegression", "CudaGPRegression", "GPBasisFuncRegression", "AbstractGPRegressionModel", "MultGaussianPRV",
    "STPRegression", "AbstractSTPRegressionModel", "MultStudentsTPRV",
    "GridSearch", "CoupledSimulatedAnnealing", "GradBasedGlobalOptimizer", "ModelTuner", "ProbGPCommMachine",
    "AdaptiveHyperParameterMCMC", "HyperParameterMCMC", "HyperParameterSCAM", "DecomposableCovariance",
    "ESGPModel", "BlockedMESNRV", "TensorCombinationKernel", "KroneckerProductKernel",
    "MOGPRegressionModel", "KroneckerMOGPModel", "CoRegRBFKernel", "CoRegCauchyKernel", "CoRegLaplaceKernel",
    "CoRegTStudentKernel", "CoRegDiracKernel",
]
