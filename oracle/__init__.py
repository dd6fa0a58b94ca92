"""CPU oracle (test infrastructure only). See closed_form.py / psd_cv2.py headers."""
