"""empty stand-in (imported at utils.py:15, unused on the hot path)"""
