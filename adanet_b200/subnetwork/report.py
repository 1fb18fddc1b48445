"""Report containers (adanet/subnetwork/report.py:29,136).  Report materialisation
(adanet/core/report_materializer.py) is outside the hot-path scope (SURVEY.md
section 2 row 14); the containers exist so Builders that override
`build_subnetwork_report` keep working and generators receive (empty) reports."""

import collections


class Report(collections.namedtuple("Report", ["hparams", "attributes", "metrics"])):
  def __new__(cls, hparams, attributes, metrics):
    return super(Report, cls).__new__(cls, hparams=dict(hparams), attributes=dict(attributes), metrics=dict(metrics))


class MaterializedReport(collections.namedtuple(
    "MaterializedReport", ["iteration_number", "name", "hparams", "attributes", "metrics", "included_in_final_ensemble"])):
  def __new__(cls, iteration_number, name, hparams, attributes, metrics, included_in_final_ensemble):
    return super(MaterializedReport, cls).__new__(cls, iteration_number, name, dict(hparams), dict(attributes),
                                                  dict(metrics), included_in_final_ensemble)
