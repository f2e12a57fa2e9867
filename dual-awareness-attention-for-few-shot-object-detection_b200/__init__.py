"""dana_b200 -- Blackwell-native (sm_100a) forward hot path of DAnA few-shot detection.

The directory name carries hyphens (it mirrors the upstream repository name), so import it through
the `dana_b200` shim at the repo root:  `import dana_b200`.
"""
__all__ = ["_lib", "ops", "build"]
