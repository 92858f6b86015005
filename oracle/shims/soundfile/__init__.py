"""empty stand-in (imported at utils.py:10, unused on the hot path)"""
