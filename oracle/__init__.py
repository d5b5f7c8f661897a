"""TEST INFRASTRUCTURE ONLY — CPU fp32 oracle of the reference hot path (see oracle/README.md)."""
