"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Never imported by easylp_b200/."""
