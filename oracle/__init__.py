"""Test-only oracles (see aru_oracle.py header). Not imported by the product."""
