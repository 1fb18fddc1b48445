"""adanet.subnetwork mirror (adanet/subnetwork/__init__.py:24-37)."""

from adanet_b200.subnetwork.generator import Builder
from adanet_b200.subnetwork.generator import Generator
from adanet_b200.subnetwork.generator import SimpleGenerator
from adanet_b200.subnetwork.generator import Subnetwork
from adanet_b200.subnetwork.generator import TrainOpSpec
from adanet_b200.subnetwork.report import MaterializedReport
from adanet_b200.subnetwork.report import Report

__all__ = ["Subnetwork", "Builder", "Generator", "SimpleGenerator", "TrainOpSpec", "Report", "MaterializedReport"]
