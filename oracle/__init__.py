"""CPU oracle (test infrastructure only; see oracle/README.md).  Never imported by the product package."""
