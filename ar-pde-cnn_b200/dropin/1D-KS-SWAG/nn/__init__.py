"""drop-in nn package (see ../../README.md)"""
