"""TEST INFRASTRUCTURE ONLY: CPU oracle of the hot path (see monkey_oracle.py).  Never imported by the product."""
