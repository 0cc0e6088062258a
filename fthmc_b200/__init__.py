"""fthmc_b200 -- B200 (sm_100a) implementation of nftqcd/fthmc's FT-HMC trajectory path.

Python here is only the host-side mirror of the reference's function interface; every body is a CUDA
kernel in libfthmc_b200.so (C ABI: include/fthmc_b200.h).  Importing the package does not need a GPU;
calling any entry point does, and raises if the library or the device is missing."""
from ._lib import FthmcError, LIB_PATH, lib  # noqa: F401
from .flow import (PackedFlow, pack, raw_weights_of, default_init_raw, raw_from_state_dict, pack_state_dict,  # noqa: F401
                   load_flow, flow_resize)
from .api import (Param, action, u1_action, force, regularize, topocharge, topo_charge, leapfrog, hmc, hmc_batch,  # noqa: F401
                  ft_flow, ft_flow_inv, ft_action, ft_force, ft_leapfrog, ft_hmc, ft_hmc_batch,
                  hmc_run_batch, ft_hmc_run_batch, run, ft_run, run_hmc, topo_history, ft_action_grad, ft_force_norm_grad, flow_vjp, differentiable_flow,
                  differentiable_u1_action)
from . import stats, shard, train, sampler  # noqa: F401
from .train import FlowTrainer, flow_train, flow_eval  # noqa: F401
from .field_transformation import FieldTransformation  # noqa: F401

__all__ = ["Param", "action", "u1_action", "force", "regularize", "topocharge", "topo_charge", "leapfrog", "hmc",
           "hmc_batch", "hmc_run_batch", "ft_hmc_run_batch", "run", "ft_run", "run_hmc", "stats", "shard", "train", "sampler", "FlowTrainer", "flow_train", "flow_eval", "FieldTransformation", "ft_action_grad", "ft_force_norm_grad", "flow_vjp", "differentiable_flow", "differentiable_u1_action", "ft_flow", "ft_flow_inv", "ft_action", "ft_force", "ft_leapfrog", "ft_hmc", "ft_hmc_batch",
           "PackedFlow", "pack", "raw_weights_of", "raw_from_state_dict", "pack_state_dict", "load_flow", "flow_resize", "FthmcError", "lib", "LIB_PATH"]
