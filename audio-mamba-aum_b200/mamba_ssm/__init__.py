"""Drop-in ``mamba_ssm`` namespace backed by the B200 engine (aum_b200).

Put this directory on PYTHONPATH ahead of (or instead of) the pip ``mamba_ssm`` wheel and the reference's
``src/models/mamba_models.py`` (:18, :26) and ``src/run.py`` import it unchanged.
Only the import paths the AuM hot path uses are provided.
"""
__version__ = "1.1.3.post1+aum_b200"
