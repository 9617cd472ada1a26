"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/fma_oracle.h).  Never imported by the product."""
