"""Shim of the detectron2 surface the reference's inference path touches (test infra)."""
