"""Constants of the hot path, same names and values as the reference's gonzag/config.py:10-37."""

ldebug = False
ivrb = 0                       # the reference defaults to 1 (chatty); the drop-in stays quiet
nb_talk = 10
deg2km = 111.11                # gonzag/config.py:20
rmissval = -9999.              # gonzag/config.py:26
rfactor = 0.75                 # gonzag/config.py:28
search_box_w_km = 500.         # gonzag/config.py:31
l_save_track_on_model_grid = True
R_eq = 6378.137
R_pl = 6356.752
