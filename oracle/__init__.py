"""oracle/ -- TEST INFRASTRUCTURE ONLY (parity checker + reported CPU baseline).

Nothing under bridgeqa_b200/ may import this package.  See oracle/README.md.
"""
