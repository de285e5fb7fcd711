"""TEST INFRASTRUCTURE ONLY: CPU oracle of the MANNeR hot path (see manner_oracle.py). Never imported by manner_b200/."""
