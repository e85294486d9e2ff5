"""CPU oracle for the PROSPECT annealing / MH hot path -- test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this
package; the product (prospect_public_b200/) never does.
"""
