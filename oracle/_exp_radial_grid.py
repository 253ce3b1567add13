# (test infrastructure: oracle copy of the same data table)
# Abscissae of the tabulated inverse CDF used by the reference's exponential walker initialisation
# (src/deeperwin/utils/utils.py:392-444, `xp`; 50 points on [0, 20]).  The ordinates `Fp` (utils.py:445-497) are the closed form
# F(x) = 1 - exp(-x) (1 + x + x^2 / 2), the CDF of p(x) = x^2 exp(-x) / 2, tabulated to 8 digits (max deviation 1.2e-9): they are
# recomputed here instead of being stored.
EXP_RADIAL_XP = (
    0.0, 0.16646017, 0.2962203, 0.42016042, 0.54028054, 0.6017006, 0.66318066, 0.78726079, 0.91426091, 1.04588105,
    1.15018115, 1.25946126, 1.37514138, 1.4998815, 1.73462173, 2.22060222, 2.43890244, 2.66892267, 2.88428288,
    3.004163, 3.12246312, 3.23980324, 3.35662336, 3.47338347, 3.59030359, 3.70764371, 3.82564383, 3.93746394,
    4.05024405, 4.16416416, 4.27938428, 4.3960844, 4.51440451, 4.63452463, 4.75662476, 5.00734501, 5.26824527,
    5.54106554, 5.82790583, 6.12646613, 6.44380644, 6.78324678, 7.14934715, 7.54514755, 7.98038798, 8.46468846,
    9.01334901, 10.13523014, 11.71389171, 20.0,
)
