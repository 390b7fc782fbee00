"""multi_part_assembly_b200 -- B200 (sm_100a) native hot path of
Wuziyi616/multi_part_assembly behind the reference's own Python API.

`import multi_part_assembly_b200.compat` (or `compat.install()`) additionally
registers this package under the name `multi_part_assembly` and provides
minimal `yacs` / `pytorch_lightning` stand-ins when those are not installed, so
the reference's config files run unmodified.
"""
__version__ = '0.1.0'
