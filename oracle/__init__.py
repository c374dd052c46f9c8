"""TEST INFRASTRUCTURE ONLY: CPU oracle for the outfit-scoring hot path (see restatement.py)."""
