"""Oracle package: CPU restatement of the reference hot path.  Test infrastructure only --
never imported by ``ufvideo_b200``."""
