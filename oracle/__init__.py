"""Test infrastructure only: CPU restatement of the reference's hot path (see sgm_oracle.py). Never imported by
ccedit_b200/."""
