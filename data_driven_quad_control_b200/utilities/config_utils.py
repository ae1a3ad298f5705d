"""YAML loading helper (reference: utilities/config_utils.py:6-8)."""

from typing import Any

import yaml


def load_yaml_config(config_file: str) -> Any:
    with open(config_file, "r") as fh:
        return yaml.safe_load(fh)
